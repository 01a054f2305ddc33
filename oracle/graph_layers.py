"""ORACLE (test infrastructure) — numpy float32 restatement of the three non-GEMM graph layers.

  proposal_layer      <- mrcnn/model.py:329-406 (ProposalLayer) with apply_box_deltas_graph :287-308,
                         clip_boxes_graph :311-326, utils.batch_slice utils.py:872-906
  pyramid_roi_align   <- mrcnn/model.py:428-534 (PyramidROIAlign), log2_graph :413-423
  detection_layer     <- mrcnn/model.py:868-909 (DetectionLayer), refine_detections_graph :770-865,
                         norm_boxes_graph :3003-3017

Third-party op semantics (tf.nn.top_k, tf.image.non_max_suppression, tf.image.crop_and_resize,
tf.round, tf.cast, tf.unique, tf.sets.set_intersection) are restated from SURVEY.md Appendix C3.
All arithmetic is float32 with one rounding per op (numpy never contracts to FMA).
"""
import numpy as np

from . import _native

F32 = np.float32


def exp_f32(x):
    """exp convention shared with the CUDA kernels: float64 evaluation, one rounding to float32."""
    return np.exp(np.asarray(x, dtype=np.float64)).astype(F32)


def log_f32(x):
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.log(np.asarray(x, dtype=np.float64)).astype(F32)


def top_k_indices(values, k):
    """tf.nn.top_k(sorted=True).indices: descending value, equal values -> lower index first."""
    values = np.asarray(values)
    # stable sort on the negated key keeps lower indices first among ties
    order = np.argsort(-values.astype(np.float64), kind="stable")
    return order[:k].astype(np.int32)


def apply_box_deltas(boxes, deltas):
    """mrcnn/model.py:287-308, float32 op by op."""
    boxes = boxes.astype(F32)
    deltas = deltas.astype(F32)
    height = boxes[:, 2] - boxes[:, 0]
    width = boxes[:, 3] - boxes[:, 1]
    center_y = boxes[:, 0] + F32(0.5) * height
    center_x = boxes[:, 1] + F32(0.5) * width
    center_y = center_y + deltas[:, 0] * height
    center_x = center_x + deltas[:, 1] * width
    height = height * exp_f32(deltas[:, 2])
    width = width * exp_f32(deltas[:, 3])
    y1 = center_y - F32(0.5) * height
    x1 = center_x - F32(0.5) * width
    y2 = y1 + height
    x2 = x1 + width
    return np.stack([y1, x1, y2, x2], axis=1).astype(F32)


def clip_boxes(boxes, window):
    """mrcnn/model.py:311-326: max(min(v, hi), lo) per coordinate."""
    wy1, wx1, wy2, wx2 = [F32(v) for v in window]
    y1 = np.maximum(np.minimum(boxes[:, 0], wy2), wy1)
    x1 = np.maximum(np.minimum(boxes[:, 1], wx2), wx1)
    y2 = np.maximum(np.minimum(boxes[:, 2], wy2), wy1)
    x2 = np.maximum(np.minimum(boxes[:, 3], wx2), wx1)
    return np.stack([y1, x1, y2, x2], axis=1).astype(F32)


def proposal_layer(rpn_class, rpn_bbox, anchors, *, pre_nms_limit=6000, proposal_count=1000,
                   nms_threshold=0.7, rpn_bbox_std_dev=(0.1, 0.1, 0.2, 0.2), return_taps=False):
    """rpn_class [B,A,2], rpn_bbox [B,A,4], anchors [A,4] or [B,A,4] -> rpn_rois [B,R,4] float32.

    With return_taps also returns per-image dicts holding the bit-exact checkpoints:
    topk (int32 [K]), boxes (f32 [K,4] decoded+clipped), keep (int32 [<=R], indices into topk order).
    """
    rpn_class = np.asarray(rpn_class, dtype=F32)
    rpn_bbox = np.asarray(rpn_bbox, dtype=F32)
    anchors = np.asarray(anchors, dtype=F32)
    B, A = rpn_class.shape[:2]
    std = np.asarray(rpn_bbox_std_dev, dtype=F32).reshape(1, 4)
    K = min(int(pre_nms_limit), A)
    out = np.zeros((B, proposal_count, 4), dtype=F32)
    taps = []
    for b in range(B):
        scores = rpn_class[b, :, 1]
        deltas = rpn_bbox[b] * std                      # model.py:355
        anc = anchors[b] if anchors.ndim == 3 else anchors
        ix = top_k_indices(scores, K)                   # model.py:361-363
        s = scores[ix]
        d = deltas[ix]
        a = anc[ix]
        boxes = apply_box_deltas(a, d)                  # model.py:374-378
        boxes = clip_boxes(boxes, (0.0, 0.0, 1.0, 1.0))  # model.py:382-386
        keep = _native.nms_tf113(boxes, s, proposal_count, nms_threshold)  # model.py:392-395
        out[b, :keep.shape[0]] = boxes[keep]            # gather + zero pad, model.py:396-400
        taps.append({"topk": ix, "boxes": boxes, "scores": s, "keep": keep})
    if return_taps:
        return out, taps
    return out


# --------------------------------------------------------------------------------------------
# PyramidROIAlign
# --------------------------------------------------------------------------------------------

def roi_levels(boxes, image_area):
    """mrcnn/model.py:465-477.  boxes [...,4] f32 normalized -> int32 level in [2,5].

    tf.log -> float64 log rounded to f32; tf.round = half-to-even; float->int32 cast truncates,
    with -inf (zero-area ROI) and NaN mapping to INT_MIN as on x86 (cvttss2si), so that
    4 + INT_MIN wraps and max(2, .) yields level 2.
    """
    boxes = np.asarray(boxes, dtype=F32)
    h = boxes[..., 2] - boxes[..., 0]
    w = boxes[..., 3] - boxes[..., 1]
    area = F32(image_area)
    with np.errstate(invalid="ignore", divide="ignore"):
        denom = F32(224.0) / np.sqrt(area, dtype=F32)
        v = np.sqrt(h * w, dtype=F32) / denom
        lv = log_f32(v) / log_f32(F32(2.0))
        r = np.rint(lv).astype(F32)                     # half-to-even
    bad = ~np.isfinite(r) | (np.abs(r) >= F32(2147483648.0))
    ri = np.where(bad, 0, r).astype(np.int64)
    ri = np.where(bad, np.int64(-2147483648), ri)
    lvl = (4 + ri)
    lvl = ((lvl + 2**31) % 2**32) - 2**31               # int32 wrap-around
    lvl = np.minimum(5, np.maximum(2, lvl))
    return lvl.astype(np.int32)


def crop_and_resize(image, boxes, box_ind, crop_size):
    """tf.image.crop_and_resize(method='bilinear', extrapolation_value=0), float32.

    image [N,H,W,C] f32, boxes [n,4] normalized (y1,x1,y2,x2), box_ind [n] -> [n,ph,pw,C].
    """
    image = np.asarray(image, dtype=F32)
    boxes = np.asarray(boxes, dtype=F32)
    n = boxes.shape[0]
    ph, pw = crop_size
    _, H, W, C = image.shape
    out = np.zeros((n, ph, pw, C), dtype=F32)
    Hm1 = F32(H - 1)
    Wm1 = F32(W - 1)
    for i in range(n):
        y1, x1, y2, x2 = boxes[i]
        img = image[box_ind[i]]
        hs = (y2 - y1) * Hm1 / F32(ph - 1) if ph > 1 else F32(0)
        ws = (x2 - x1) * Wm1 / F32(pw - 1) if pw > 1 else F32(0)
        for iy in range(ph):
            in_y = y1 * Hm1 + F32(iy) * hs if ph > 1 else F32(0.5) * (y1 + y2) * Hm1
            if not (in_y >= 0 and in_y <= Hm1):          # also rejects NaN
                continue
            t = int(np.floor(in_y))
            bt = int(np.ceil(in_y))
            ly = F32(in_y - F32(t))
            for ix in range(pw):
                in_x = x1 * Wm1 + F32(ix) * ws if pw > 1 else F32(0.5) * (x1 + x2) * Wm1
                if not (in_x >= 0 and in_x <= Wm1):
                    continue
                l = int(np.floor(in_x))
                r = int(np.ceil(in_x))
                lx = F32(in_x - F32(l))
                tl = img[t, l]
                tr = img[t, r]
                bl = img[bt, l]
                br = img[bt, r]
                top = tl + (tr - tl) * lx
                bot = bl + (br - bl) * lx
                out[i, iy, ix] = top + (bot - top) * ly
    return out


def crop_and_resize_fast(image, boxes, box_ind, crop_size):
    """Vectorised twin of crop_and_resize (same op order, same float32 roundings)."""
    image = np.asarray(image, dtype=F32)
    boxes = np.asarray(boxes, dtype=F32)
    box_ind = np.asarray(box_ind)
    n = boxes.shape[0]
    ph, pw = crop_size
    _, H, W, C = image.shape
    Hm1 = F32(H - 1)
    Wm1 = F32(W - 1)
    y1, x1, y2, x2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    if ph > 1:
        hs = (y2 - y1) * Hm1 / F32(ph - 1)
        in_y = (y1 * Hm1)[:, None] + np.arange(ph, dtype=F32)[None, :] * hs[:, None]
    else:
        in_y = (F32(0.5) * (y1 + y2) * Hm1)[:, None]
    if pw > 1:
        ws = (x2 - x1) * Wm1 / F32(pw - 1)
        in_x = (x1 * Wm1)[:, None] + np.arange(pw, dtype=F32)[None, :] * ws[:, None]
    else:
        in_x = (F32(0.5) * (x1 + x2) * Wm1)[:, None]
    in_y = in_y.astype(F32)
    in_x = in_x.astype(F32)
    with np.errstate(invalid="ignore"):
        vy = (in_y >= 0) & (in_y <= Hm1)
        vx = (in_x >= 0) & (in_x <= Wm1)
    sy = np.where(vy, in_y, F32(0))
    sx = np.where(vx, in_x, F32(0))
    t = np.floor(sy).astype(np.int64)
    bt = np.ceil(sy).astype(np.int64)
    l = np.floor(sx).astype(np.int64)
    r = np.ceil(sx).astype(np.int64)
    ly = (sy - t.astype(F32)).astype(F32)[:, :, None, None]
    lx = (sx - l.astype(F32)).astype(F32)[:, None, :, None]
    bi = box_ind[:, None, None]
    tl = image[bi, t[:, :, None], l[:, None, :]]
    tr = image[bi, t[:, :, None], r[:, None, :]]
    bl = image[bi, bt[:, :, None], l[:, None, :]]
    br = image[bi, bt[:, :, None], r[:, None, :]]
    top = tl + (tr - tl) * lx
    bot = bl + (br - bl) * lx
    out = top + (bot - top) * ly
    valid = (vy[:, :, None] & vx[:, None, :])[..., None]
    return np.where(valid, out, F32(0)).astype(F32)


def pyramid_roi_align(boxes, image_shape, feature_maps, pool_shape, return_levels=False):
    """boxes [B,N,4] f32; feature_maps = [P2,P3,P4,P5] each [B,Hl,Wl,C] f32 -> [B,N,p,p,C].

    The reference routes per level, crops, concatenates and then restores (batch, box) order
    (model.py:480-531); the net effect is out[b,n] = crop(P_level(b,n)[b], boxes[b,n]).
    """
    boxes = np.asarray(boxes, dtype=F32)
    B, N = boxes.shape[:2]
    C = feature_maps[0].shape[-1]
    ph, pw = pool_shape
    area = F32(F32(image_shape[0]) * F32(image_shape[1]))
    lvls = roi_levels(boxes, area)
    out = np.zeros((B, N, ph, pw, C), dtype=F32)
    for i, level in enumerate(range(2, 6)):
        bsel, nsel = np.where(lvls == level)
        if bsel.size == 0:
            continue
        out[bsel, nsel] = crop_and_resize_fast(feature_maps[i], boxes[bsel, nsel], bsel, (ph, pw))
    if return_levels:
        return out, lvls
    return out


# --------------------------------------------------------------------------------------------
# DetectionLayer
# --------------------------------------------------------------------------------------------

def norm_window(window_px, image_shape_hw):
    """norm_boxes_graph, model.py:3003-3017: (box - [0,0,1,1]) / ([h,w,h,w] - 1) in float32."""
    h, w = F32(image_shape_hw[0]), F32(image_shape_hw[1])
    scale = np.array([h, w, h, w], dtype=F32) - F32(1.0)
    shift = np.array([0.0, 0.0, 1.0, 1.0], dtype=F32)
    return ((np.asarray(window_px, dtype=F32) - shift) / scale).astype(F32)


def refine_detections(rois, probs, deltas, window, *, bbox_std_dev=(0.1, 0.1, 0.2, 0.2),
                      min_confidence=0.0, nms_threshold=0.3, max_instances=100,
                      return_taps=False):
    """mrcnn/model.py:770-865 for one image. rois [N,4], probs [N,NC], deltas [N,NC,4]."""
    rois = np.asarray(rois, dtype=F32)
    probs = np.asarray(probs, dtype=F32)
    deltas = np.asarray(deltas, dtype=F32)
    N = probs.shape[0]
    class_ids = np.argmax(probs, axis=1).astype(np.int32)          # first maximal index
    ar = np.arange(N)
    class_scores = probs[ar, class_ids]
    deltas_specific = deltas[ar, class_ids] * np.asarray(bbox_std_dev, dtype=F32).reshape(1, 4)
    refined = apply_box_deltas(rois, deltas_specific)
    refined = clip_boxes(refined, window)
    keep = np.where(class_ids > 0)[0]
    if min_confidence:                                            # truthiness, model.py:804
        conf_keep = np.where(class_scores >= F32(min_confidence))[0]
        keep = np.array(sorted(set(keep.tolist()) & set(conf_keep.tolist())), dtype=np.int64)
    pre_cls = class_ids[keep]
    pre_scores = class_scores[keep]
    pre_rois = refined[keep]
    # tf.unique keeps first-occurrence order
    _, first = np.unique(pre_cls, return_index=True)
    uniq = pre_cls[np.sort(first)]
    nms_keep = []
    for cid in uniq:
        ixs = np.where(pre_cls == cid)[0]
        ck = _native.nms_tf113(pre_rois[ixs], pre_scores[ixs], max_instances, nms_threshold)
        nms_keep.extend(keep[ixs[ck]].tolist())
    # set_intersection -> ascending sorted unique values
    keep2 = np.array(sorted(set(keep.tolist()) & set(nms_keep)), dtype=np.int64)
    scores_keep = class_scores[keep2]
    num_keep = min(scores_keep.shape[0], max_instances)
    top_ids = top_k_indices(scores_keep, num_keep)
    final = keep2[top_ids]
    det = np.zeros((max_instances, 6), dtype=F32)
    det[:final.shape[0], :4] = refined[final]
    det[:final.shape[0], 4] = class_ids[final].astype(F32)
    det[:final.shape[0], 5] = class_scores[final]
    if return_taps:
        return det, {"class_ids": class_ids, "class_scores": class_scores, "refined": refined,
                     "final": final.astype(np.int32)}
    return det


def detection_layer(rois, mrcnn_class, mrcnn_bbox, image_meta, **kw):
    """mrcnn/model.py:881-906. image_meta [B, 12+NC]; window normalised with image_shape of
    the first image (model.py:893-895)."""
    image_meta = np.asarray(image_meta, dtype=F32)
    image_shape = image_meta[0, 4:7]
    B = rois.shape[0]
    out = []
    for b in range(B):
        window = norm_window(image_meta[b, 7:11], image_shape[:2])
        out.append(refine_detections(rois[b], mrcnn_class[b], mrcnn_bbox[b], window, **kw))
    return np.stack(out).astype(F32)
