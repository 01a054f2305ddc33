"""mrcnn.utils — host-side mirror of the utilities the detect path uses (reference:
mrcnn/utils.py).  Array maths that sits on the hot path (zscale stretch, uint8 conversion, resize /
pad) runs on the GPU through libmrcnn_b200.so; there is no CPU fallback.  Anchor generation and box
normalisation are cached one-off host computations, as in the reference.

Mirrors: read_fits :1033-1163 (default flag set), get_fits_header :992-1005, get_fits_size
:1009-1030, resize_image :456-561 (modes none/square), compute/generate anchors :652-708,
norm_boxes :923-937, denorm_boxes :940-954.
"""
import logging
import math

import numpy as np

from . import _native, fitsio

logger = logging.getLogger("mrcnn")


# --------------------------------------------------------------------------------------------
# device helpers
# --------------------------------------------------------------------------------------------

def _torch():
    import torch
    if not torch.cuda.is_available():
        raise _native.NativeError("mrcnn (B200 build) needs a CUDA device; there is no CPU fallback")
    return torch


def maps_to_rgb8_device(maps, zscale_contrasts=(0.25, 0.25, 0.25)):
    """maps: CUDA float32 tensor [n,H,W] (NaN allowed) -> (rgb uint8 [n,H,W,3], minmax int32 [n,2]).
    The GPU form of read_fits' NaN fill + zscale + normalise + uint8 RGB (utils.py:1090-1208)."""
    torch = _torch()
    lib = _native.lib()
    assert maps.is_cuda and maps.dtype == torch.float32 and maps.dim() == 3 and maps.is_contiguous()
    n, H, W = maps.shape
    params = torch.empty((n, 3, 4), dtype=torch.float32, device=maps.device)
    rgb = torch.empty((n, H, W, 3), dtype=torch.uint8, device=maps.device)
    minmax = torch.empty((n, 2), dtype=torch.int32, device=maps.device)
    con = _native.float_array(list(zscale_contrasts))
    with torch.cuda.device(maps.device):        # launch on the tensor's device, whatever the process default is
        stream = torch.cuda.current_stream(maps.device).cuda_stream
        _native.check(lib.mrcnn_zscale_params(_native.ptr(maps), n, H, W, con, _native.ptr(params), stream), "zscale_params")
        _native.check(lib.mrcnn_stretch_to_rgb8(_native.ptr(maps), _native.ptr(params), n, H, W, _native.ptr(rgb),
                                                _native.ptr(minmax), stream), "stretch_to_rgb8")
    return rgb, minmax, params


def square_geometry(h, w, min_dim, max_dim, min_scale, mode):
    """scale / output size / padding / window of utils.resize_image (utils.py:489-533)."""
    scale = 1
    if mode == "none":
        return 1, (h, w), (0, 0), (0, 0, h, w), [(0, 0), (0, 0), (0, 0)]
    if mode != "square":
        raise Exception("Mode {} not supported".format(mode))
    if min_dim:
        scale = max(1, min_dim / min(h, w))
    if min_scale and scale < min_scale:
        scale = min_scale
    if max_dim:
        image_max = max(h, w)
        if round(image_max * scale) > max_dim:
            scale = max_dim / image_max
    oh, ow = (round(h * scale), round(w * scale)) if scale != 1 else (h, w)
    top = (max_dim - oh) // 2
    left = (max_dim - ow) // 2
    padding = [(top, max_dim - oh - top), (left, max_dim - ow - left), (0, 0)]
    window = (top, left, oh + top, ow + left)
    return scale, (oh, ow), (top, left), window, padding


def mold_rgb8_device(rgb, minmax, out_hw, square, top_left, mean_pixel, out=None):
    """rgb CUDA uint8 [n,H,W,3] -> molded CUDA float32 [n,S,S,3] (resize + pad + mean subtraction)."""
    torch = _torch()
    lib = _native.lib()
    n, H, W, _ = rgb.shape
    if minmax is None:
        flat = rgb.reshape(n, -1)
        minmax = torch.stack([flat.amin(dim=1), flat.amax(dim=1)], dim=1).to(torch.int32).contiguous()
    if out is None:
        out = torch.empty((n, square, square, 3), dtype=torch.float32, device=rgb.device)
    mean = _native.float_array([float(v) for v in np.asarray(mean_pixel).reshape(-1)[:3]])
    with torch.cuda.device(rgb.device):
        stream = torch.cuda.current_stream(rgb.device).cuda_stream
        _native.check(lib.mrcnn_resize_pad_mold(_native.ptr(rgb), _native.ptr(minmax), n, H, W, int(out_hw[0]), int(out_hw[1]),
                                                int(square), int(top_left[0]), int(top_left[1]), mean, _native.ptr(out), stream),
                      "resize_pad_mold")
    return out


# --------------------------------------------------------------------------------------------
# FITS
# --------------------------------------------------------------------------------------------

def get_fits_header(filename):
    try:
        header, _ = fitsio.read_header_only(filename)         # header blocks only: survey images are GBs
    except Exception:
        logger.error("ERROR: Cannot read image file: " + str(filename))
        return None
    return header


def get_fits_size(filename):
    header = get_fits_header(filename)
    if header is None:
        return None
    for key in ("NAXIS1", "NAXIS2"):
        if key not in header:
            logger.error("%s keyword missing in header!" % key)
            return None
    return header["NAXIS1"], header["NAXIS2"]


def read_fits(filename, xmin=-1, xmax=-1, ymin=-1, ymax=-1, stretch=True, normalize=True, convertToRGB=True,
              zscale_contrasts=[0.25, 0.25, 0.25], to_uint8=True, stretch_biascontrast=False, contrast=1, bias=0.5):
    """FITS file -> ([H,W,3] uint8 image, header); None when the file cannot be read or is not a
    2-D / 4-D image (same convention as the reference).  Only the flag combination `run.py detect`
    uses is supported (zscale stretch + normalise + uint8 RGB); other combinations raise."""
    if len(zscale_contrasts) != 3:
        logger.warning("Size of input zscale_contrasts is !=3, ignoring inputs and using default (0.25,0.25,0.25)...")
        zscale_contrasts = [0.25, 0.25, 0.25]
    try:
        data, header = fitsio.read_primary(filename)
    except Exception:
        logger.error("ERROR: Cannot read image file: " + str(filename))
        return None
    read_tile = (xmin >= 0 and xmax >= 0 and ymin >= 0 and ymax >= 0)
    if read_tile:
        if xmax <= xmin:
            logger.error("xmax must be >xmin for tile reading!")
            return None
        if ymax <= ymin:
            logger.error("ymax must be >ymin for tile reading!")
            return None
    if data is None or data.ndim not in (2, 4):
        logger.error("ERROR: Invalid/unsupported number of channels found in file " + str(filename))
        return None
    plane = data[0, 0] if data.ndim == 4 else data
    if read_tile:
        plane = plane[ymin:ymax, xmin:xmax]
    if not stretch and not normalize and not convertToRGB and not stretch_biascontrast:
        # the raw plane (ground-truth mask files of the training datasets, scripts/run.py:641-727): float32, NaN -> minimum
        out = np.array(plane, dtype=np.float32)
        bad = np.isnan(out)
        if bad.any():
            out[bad] = np.nanmin(out)
        return out, header
    if not (stretch and normalize and convertToRGB and to_uint8) or stretch_biascontrast:
        raise NotImplementedError("read_fits (B200 build): only stretch=True, normalize=True, convertToRGB=True, "
                                  "to_uint8=True, stretch_biascontrast=False is implemented (SURVEY.md §8a row a17)")
    torch = _torch()
    maps = torch.from_numpy(np.ascontiguousarray(plane, dtype=np.float32)).cuda().unsqueeze(0).contiguous()
    rgb, _, _ = maps_to_rgb8_device(maps, zscale_contrasts)
    return rgb[0].cpu().numpy(), header


# --------------------------------------------------------------------------------------------
# resize (host API, GPU arithmetic)
# --------------------------------------------------------------------------------------------

def resize_image(image, min_dim=None, max_dim=None, min_scale=None, mode="square"):
    """uint8 [H,W,3] -> (image, window, scale, padding, crop), reference utils.resize_image."""
    h, w = image.shape[:2]
    if mode == "none":
        return image, (0, 0, h, w), 1, [(0, 0), (0, 0), (0, 0)], None
    if image.dtype != np.uint8 or image.ndim != 3 or image.shape[2] != 3:
        raise NotImplementedError("resize_image (B200 build): uint8 [H,W,3] images only")
    scale, out_hw, top_left, window, padding = square_geometry(h, w, min_dim, max_dim, min_scale, mode)
    if scale == 1:          # no resampling (reference utils.py:514): zero padding only, on the host like the reference
        return np.pad(image, padding, mode='constant', constant_values=0), window, scale, padding, None
    torch = _torch()
    rgb = torch.from_numpy(np.ascontiguousarray(image)).cuda().unsqueeze(0)
    molded = mold_rgb8_device(rgb, None, out_hw, max_dim, top_left, (0.0, 0.0, 0.0))
    return molded[0].to(torch.uint8).cpu().numpy(), window, scale, padding, None


def resize(image, output_shape, order=1, mode='constant', cval=0, clip=True, preserve_range=False, anti_aliasing=False,
           anti_aliasing_sigma=None):
    """reference: mrcnn/utils.py:957-978 — the wrapper around skimage.transform.resize; here the scikit-image <= 0.15
    bilinear warp runs on the GPU (float64, same operation order). Only the argument values the reference itself uses
    are implemented (order=1, mode='constant', cval=0, clip=True, no anti-aliasing). Returns float64 like skimage:
    float inputs keep their range, uint8 / bool inputs are scaled to [0,1] unless preserve_range."""
    if order != 1 or mode != 'constant' or cval != 0 or not clip or anti_aliasing:
        raise NotImplementedError("resize (B200 build): only order=1, mode='constant', cval=0, clip=True, anti_aliasing=False")
    image = np.asarray(image)
    if preserve_range or image.dtype.kind == "f" or image.dtype == np.bool_:
        img = image.astype(np.float64)
    elif image.dtype == np.uint8:
        img = image.astype(np.float64) / 255.0
    else:
        raise NotImplementedError("resize (B200 build): dtype %s" % image.dtype)
    rows, cols = int(output_shape[0]), int(output_shape[1])
    squeeze = img.ndim == 2
    if squeeze:
        img = img[:, :, None]
    if img.ndim != 3:
        raise NotImplementedError("resize (B200 build): [H,W] or [H,W,C] images only")
    if rows == 0 or cols == 0 or img.size == 0:
        out = np.zeros((rows, cols, img.shape[2]), dtype=np.float64)
        return out[:, :, 0] if squeeze else out
    torch = _torch()
    lib = _native.lib()
    d_in = torch.from_numpy(np.ascontiguousarray(img)).cuda()
    d_out = torch.empty((rows, cols, img.shape[2]), dtype=torch.float64, device=d_in.device)
    with torch.cuda.device(d_in.device):
        _native.check(lib.mrcnn_skimage_resize_f64(_native.ptr(d_in), img.shape[0], img.shape[1], img.shape[2], rows, cols,
                                                   float(img.min()), float(img.max()), _native.ptr(d_out),
                                                   torch.cuda.current_stream(d_in.device).cuda_stream), "skimage_resize")
    out = d_out.cpu().numpy()
    return out[:, :, 0] if squeeze else out


def unmold_mask(mask, bbox, image_shape):
    """reference: mrcnn/utils.py:629-645 — one [h,w] float mask (28x28) resized to its box, thresholded at 0.5 and pasted
    into a full-size boolean image. (MaskRCNN.detect does this for a whole batch in one kernel, csrc/unmold.cu.)"""
    threshold = 0.5
    y1, x1, y2, x2 = bbox
    mask = resize(mask, (y2 - y1, x2 - x1))
    mask = np.where(mask >= threshold, 1, 0).astype(bool)
    full_mask = np.zeros(image_shape[:2], dtype=bool)
    full_mask[y1:y2, x1:x2] = mask
    return full_mask


def extract_bboxes(mask):
    """reference: mrcnn/utils.py:49-77 — mask [H,W,N] (0/1) -> int32 [N,(y1,x1,y2,x2)], y2 / x2 exclusive, zeros for an
    empty mask. Host numpy like the reference (the Analyzer path gets its boxes from mrcnn_planes_area_bbox on the
    device-resident bit-planes instead)."""
    mask = np.asarray(mask)
    n = mask.shape[-1]
    boxes = np.zeros([n, 4], dtype=np.int32)
    if n == 0 or mask.shape[0] == 0 or mask.shape[1] == 0:
        return boxes
    m = mask.astype(bool)
    cols = m.any(axis=0)          # [W,N]
    rows = m.any(axis=1)          # [H,N]
    has = cols.any(axis=0)
    x1 = cols.argmax(axis=0)
    x2 = cols.shape[0] - cols[::-1].argmax(axis=0)
    y1 = rows.argmax(axis=0)
    y2 = rows.shape[0] - rows[::-1].argmax(axis=0)
    boxes[has] = np.stack([y1, x1, y2, x2], axis=1)[has]
    return boxes


# --------------------------------------------------------------------------------------------
# training-path host helpers (reference: mrcnn/utils.py:147-163, 275-298, 564-622, 715-722)
# --------------------------------------------------------------------------------------------

def compute_overlaps(boxes1, boxes2):
    """IoU matrix [len(boxes1), len(boxes2)] of pixel boxes (y1,x1,y2,x2); reference utils.compute_overlaps / compute_iou
    (:75-97, :147-163): float64 arithmetic on whatever dtype comes in, intersection clipped at 0, no epsilon."""
    b1 = np.asarray(boxes1)
    b2 = np.asarray(boxes2)
    area1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    area2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    y1 = np.maximum(b1[:, None, 0], b2[None, :, 0])
    y2 = np.minimum(b1[:, None, 2], b2[None, :, 2])
    x1 = np.maximum(b1[:, None, 1], b2[None, :, 1])
    x2 = np.minimum(b1[:, None, 3], b2[None, :, 3])
    inter = np.maximum(x2 - x1, 0) * np.maximum(y2 - y1, 0)
    union = area1[:, None] + area2[None, :] - inter
    with np.errstate(divide="ignore", invalid="ignore"):
        return (inter / union).astype(np.float64)


def compute_iou(box, boxes, box_area, boxes_area):
    """IoU of one box (y1,x1,y2,x2) with each row of `boxes`; the areas come from the caller (reference utils.py:75-93)."""
    boxes = np.asarray(boxes)
    dx = np.maximum(np.minimum(box[3], boxes[:, 3]) - np.maximum(box[1], boxes[:, 1]), 0)
    dy = np.maximum(np.minimum(box[2], boxes[:, 2]) - np.maximum(box[0], boxes[:, 0]), 0)
    inter = dx * dy
    return inter / (box_area + boxes_area - inter)


def get_iou(bb1, bb2):
    """IoU of two (y1, x1, y2, x2) boxes with positive extent, 0.0 when they do not overlap (reference utils.py:100-144,
    used by the Analyzer's ground-truth comparison)."""
    ya, xa, yb, xb = bb1[0], bb1[1], bb1[2], bb1[3]
    yc, xc, yd, xd = bb2[0], bb2[1], bb2[2], bb2[3]
    assert xa < xb and ya < yb and xc < xd and yc < yd
    left, top, right, bottom = max(xa, xc), max(ya, yc), min(xb, xd), min(yb, yd)
    if right < left or bottom < top:
        return 0.0
    inter = (right - left) * (bottom - top)
    iou = inter / float((xb - xa) * (yb - ya) + (xd - xc) * (yd - yc) - inter)
    assert 0.0 <= iou <= 1.0
    return iou


def compute_overlaps_masks(masks1, masks2):
    """IoU matrix [n1, n2] of two mask stacks [H, W, n] thresholded at 0.5 (reference utils.py:166-185): float32 areas and
    intersections through one matrix product."""
    n1, n2 = masks1.shape[-1], masks2.shape[-1]
    if n1 == 0 or n2 == 0:
        return np.zeros((n1, n2))
    a = (masks1 > .5).reshape(-1, n1).astype(np.float32)
    b = (masks2 > .5).reshape(-1, n2).astype(np.float32)
    inter = a.T @ b
    return inter / (a.sum(axis=0)[:, None] + b.sum(axis=0)[None, :] - inter)


def non_max_suppression(boxes, scores, threshold):
    """Greedy NMS on the host -> int32 indices of the kept boxes, best score first (reference utils.py:188-222): candidates in
    `scores.argsort()[::-1]` order, a kept box removes every later candidate whose IoU with it exceeds `threshold`."""
    assert boxes.shape[0] > 0
    if boxes.dtype.kind != "f":
        boxes = boxes.astype(np.float32)
    area = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])
    order = scores.argsort()[::-1]
    alive = np.ones(len(order), dtype=bool)
    kept = []
    for pos in range(len(order)):
        if not alive[pos]:
            continue
        i = order[pos]
        kept.append(i)
        later = np.nonzero(alive[pos + 1:])[0] + pos + 1
        if len(later):
            rest = order[later]
            alive[later[compute_iou(boxes[i], boxes[rest], area[i], area[rest]) > threshold]] = False
    return np.array(kept, dtype=np.int32)


def apply_box_deltas(boxes, deltas):
    """Boxes (y1,x1,y2,x2) moved by (dy, dx, log dh, log dw) -> float32 [N,4] (reference utils.py:225-246)."""
    boxes = boxes.astype(np.float32)
    h = boxes[:, 2] - boxes[:, 0]
    w = boxes[:, 3] - boxes[:, 1]
    cy = boxes[:, 0] + 0.5 * h
    cx = boxes[:, 1] + 0.5 * w
    cy += deltas[:, 0] * h
    cx += deltas[:, 1] * w
    h *= np.exp(deltas[:, 2])
    w *= np.exp(deltas[:, 3])
    y1 = cy - 0.5 * h
    x1 = cx - 0.5 * w
    return np.stack([y1, x1, y1 + h, x1 + w], axis=1)


def compute_matches(gt_boxes, gt_class_ids, gt_masks, pred_boxes, pred_class_ids, pred_scores, pred_masks,
                    iou_threshold=0.5, score_threshold=0.0):
    """Greedy matching of predictions (best score first) to ground-truth instances by mask IoU and class (reference
    utils.py:725-781) -> (gt_match [n_gt], pred_match [n_pred] in score order, overlaps [n_pred, n_gt]); -1 = unmatched."""
    gt_boxes = trim_zeros(gt_boxes)
    gt_masks = gt_masks[..., :gt_boxes.shape[0]]
    pred_boxes = trim_zeros(pred_boxes)
    pred_scores = pred_scores[:pred_boxes.shape[0]]
    by_score = np.argsort(pred_scores)[::-1]
    pred_boxes, pred_class_ids, pred_masks = pred_boxes[by_score], pred_class_ids[by_score], pred_masks[..., by_score]
    overlaps = compute_overlaps_masks(pred_masks, gt_masks)
    pred_match = np.full([pred_boxes.shape[0]], -1.0)
    gt_match = np.full([gt_boxes.shape[0]], -1.0)
    for i in range(len(pred_boxes)):
        candidates = np.argsort(overlaps[i])[::-1]
        too_low = np.nonzero(overlaps[i, candidates] < score_threshold)[0]
        if too_low.size:
            candidates = candidates[:too_low[0]]
        for j in candidates:
            if gt_match[j] > -1:
                continue                      # taken by a better-scoring prediction
            if overlaps[i, j] < iou_threshold:
                break                         # candidates come in descending IoU order
            if pred_class_ids[i] == gt_class_ids[j]:
                gt_match[j], pred_match[i] = i, j
                break
    return gt_match, pred_match, overlaps


def compute_ap(gt_boxes, gt_class_ids, gt_masks, pred_boxes, pred_class_ids, pred_scores, pred_masks, iou_threshold=0.5):
    """VOC-style average precision at one IoU threshold (reference utils.py:784-820) -> (mAP, precisions, recalls, overlaps)."""
    gt_match, pred_match, overlaps = compute_matches(gt_boxes, gt_class_ids, gt_masks, pred_boxes, pred_class_ids, pred_scores,
                                                     pred_masks, iou_threshold)
    hits = np.cumsum(pred_match > -1)
    precisions = np.concatenate([[0], hits / (np.arange(len(pred_match)) + 1), [0]])
    recalls = np.concatenate([[0], hits.astype(np.float32) / len(gt_match), [1]])
    precisions = np.maximum.accumulate(precisions[::-1])[::-1]         # the best precision at this or any later recall
    step = np.nonzero(recalls[:-1] != recalls[1:])[0] + 1
    mAP = np.sum((recalls[step] - recalls[step - 1]) * precisions[step])
    return mAP, precisions, recalls, overlaps


def compute_ap_range(gt_box, gt_class_id, gt_mask, pred_box, pred_class_id, pred_score, pred_mask, iou_thresholds=None, verbose=1):
    """Mean of compute_ap over IoU thresholds, 0.5 ... 0.95 in steps of 0.05 by default (reference utils.py:823-844)."""
    iou_thresholds = iou_thresholds or np.arange(0.5, 1.0, 0.05)
    aps = []
    for thr in iou_thresholds:
        ap = compute_ap(gt_box, gt_class_id, gt_mask, pred_box, pred_class_id, pred_score, pred_mask, iou_threshold=thr)[0]
        if verbose:
            print("AP @{:.2f}:\t {:.3f}".format(thr, ap))
        aps.append(ap)
    mean_ap = np.array(aps).mean()
    if verbose:
        print("AP @{:.2f}-{:.2f}:\t {:.3f}".format(iou_thresholds[0], iou_thresholds[-1], mean_ap))
    return mean_ap


def compute_recall(pred_boxes, gt_boxes, iou):
    """Fraction of ground-truth boxes that are the best match (box IoU >= iou) of some prediction (reference utils.py:847-863)
    -> (recall, indices of those predictions)."""
    overlaps = compute_overlaps(pred_boxes, gt_boxes)
    best = np.argmax(overlaps, axis=1)
    positive_ids = np.nonzero(np.max(overlaps, axis=1) >= iou)[0]
    return len(set(best[positive_ids])) / gt_boxes.shape[0], positive_ids


def box_refinement(box, gt_box):
    """(dy, dx, log dh, log dw) that turns box into gt_box, float32 like the reference (utils.py:275-298)."""
    box = np.asarray(box).astype(np.float32)
    gt = np.asarray(gt_box).astype(np.float32)
    h, w = box[:, 2] - box[:, 0], box[:, 3] - box[:, 1]
    cy, cx = box[:, 0] + 0.5 * h, box[:, 1] + 0.5 * w
    gh, gw = gt[:, 2] - gt[:, 0], gt[:, 3] - gt[:, 1]
    gcy, gcx = gt[:, 0] + 0.5 * gh, gt[:, 1] + 0.5 * gw
    return np.stack([(gcy - cy) / h, (gcx - cx) / w, np.log(gh / h), np.log(gw / w)], axis=1)


def resize_mask(mask, scale, padding, crop=None):
    """Instance masks [H,W,N] follow their image through resize_image: nearest-neighbour zoom (scipy, order 0, as the
    reference calls it, utils.py:564-583), then the same crop or zero padding."""
    import warnings
    import scipy.ndimage
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mask = scipy.ndimage.zoom(mask, zoom=[scale, scale, 1], order=0)
    if crop is not None:
        y, x, h, w = crop
        return mask[y:y + h, x:x + w]
    return np.pad(mask, padding, mode='constant', constant_values=0)


def minimize_mask(bbox, mask, mini_shape):
    """Each instance mask cropped to its box and resized (bilinear, rounded) to mini_shape (utils.py:586-603)."""
    mini = np.zeros(tuple(mini_shape) + (mask.shape[-1],), dtype=bool)
    for i in range(mask.shape[-1]):
        y1, x1, y2, x2 = bbox[i][:4]
        m = mask[:, :, i].astype(bool)[y1:y2, x1:x2]
        if m.size == 0:
            raise Exception("Invalid bounding box with area of zero")
        mini[:, :, i] = np.around(resize(m, mini_shape)).astype(bool)
    return mini


def expand_mask(bbox, mini_mask, image_shape):
    """Inverse of minimize_mask (utils.py:606-622)."""
    mask = np.zeros(tuple(image_shape[:2]) + (mini_mask.shape[-1],), dtype=bool)
    for i in range(mask.shape[-1]):
        y1, x1, y2, x2 = bbox[i][:4]
        mask[y1:y2, x1:x2, i] = np.around(resize(mini_mask[:, :, i], (y2 - y1, x2 - x1))).astype(bool)
    return mask


def trim_zeros(x):
    """Rows that are not all zero (utils.py:715-722)."""
    assert len(x.shape) == 2
    return x[~np.all(x == 0, axis=1)]


class Dataset(object):
    """reference: mrcnn/utils.py:305-440 — the dataset bookkeeping base class (class / image registries and id maps).
    Subclasses provide load_image / load_mask; nothing here touches the GPU."""

    def __init__(self, class_map=None):
        self._image_ids = []
        self.image_info = []
        self.class_info = [{"source": "", "id": 0, "name": "BG"}]      # background is always class 0
        self.source_class_ids = {}

    def add_class(self, source, class_id, class_name):
        assert "." not in source, "Source name cannot contain a dot"
        if any(info["source"] == source and info["id"] == class_id for info in self.class_info):
            return
        self.class_info.append({"source": source, "id": class_id, "name": class_name})

    def add_image(self, source, image_id, path, **kwargs):
        info = {"id": image_id, "source": source, "path": path}
        info.update(kwargs)
        self.image_info.append(info)

    def image_reference(self, image_id):
        return ""

    def prepare(self, class_map=None):
        self.num_classes = len(self.class_info)
        self.class_ids = np.arange(self.num_classes)
        self.class_names = [",".join(c["name"].split(",")[:1]) for c in self.class_info]
        self.num_images = len(self.image_info)
        self._image_ids = np.arange(self.num_images)
        self.class_from_source_map = {"{}.{}".format(info["source"], info["id"]): i
                                      for info, i in zip(self.class_info, self.class_ids)}
        self.image_from_source_map = {"{}.{}".format(info["source"], info["id"]): i
                                      for info, i in zip(self.image_info, self.image_ids)}
        self.sources = list(set(i["source"] for i in self.class_info))
        self.source_class_ids = {}
        for source in self.sources:
            self.source_class_ids[source] = [i for i, info in enumerate(self.class_info) if i == 0 or source == info["source"]]

    def map_source_class_id(self, source_class_id):
        return self.class_from_source_map[source_class_id]

    def get_source_class_id(self, class_id, source):
        info = self.class_info[class_id]
        assert info["source"] == source
        return info["id"]

    @property
    def image_ids(self):
        return self._image_ids

    def source_image_link(self, image_id):
        return self.image_info[image_id]["path"]

    def load_image(self, image_id):
        """[H,W,3] array of the image. The reference reads any raster file through skimage.io; this build reads FITS
        (the data format of the detect path) through read_fits."""
        path = self.image_info[image_id]["path"]
        res = read_fits(path)
        if res is None:
            raise IOError("cannot read image " + str(path))
        return res[0]

    def load_mask(self, image_id):
        logging.warning("You are using the default load_mask(), maybe you need to define your own one.")
        return np.empty([0, 0, 0]), np.empty([0], np.int32)


# --------------------------------------------------------------------------------------------
# anchors / box normalisation (host, cached by the model)
# --------------------------------------------------------------------------------------------

def compute_backbone_shapes(config, image_shape):
    if callable(config.BACKBONE):
        return config.COMPUTE_BACKBONE_SHAPE(image_shape)
    assert config.BACKBONE in ["resnet50", "resnet101", "custom"]
    return np.array([[int(math.ceil(image_shape[0] / s)), int(math.ceil(image_shape[1] / s))]
                     for s in config.BACKBONE_STRIDES])


def generate_anchors(scales, ratios, shape, feature_stride, anchor_stride):
    """[(y,x) grid positions x ratios, (y1,x1,y2,x2)] anchors of one pyramid level, float64."""
    scales = np.atleast_1d(np.asarray(scales))
    ratios = np.asarray(ratios)
    # ratio-major over scales, as np.meshgrid(scales, ratios).flatten() orders them
    sc = np.repeat(scales[None, :], len(ratios), axis=0).reshape(-1)
    ra = np.repeat(ratios[:, None], len(scales), axis=1).reshape(-1)
    hh = sc / np.sqrt(ra)
    ww = sc * np.sqrt(ra)
    ys = np.arange(0, shape[0], anchor_stride) * feature_stride
    xs = np.arange(0, shape[1], anchor_stride) * feature_stride
    cy = np.broadcast_to(ys[:, None, None], (len(ys), len(xs), len(hh)))
    cx = np.broadcast_to(xs[None, :, None], (len(ys), len(xs), len(hh)))
    h = np.broadcast_to(hh[None, None, :], cy.shape)
    w = np.broadcast_to(ww[None, None, :], cy.shape)
    boxes = np.stack([cy - 0.5 * h, cx - 0.5 * w, cy + 0.5 * h, cx + 0.5 * w], axis=-1)
    return boxes.reshape(-1, 4)


def generate_pyramid_anchors(scales, ratios, feature_shapes, feature_strides, anchor_stride):
    return np.concatenate([generate_anchors(scales[i], ratios, feature_shapes[i], feature_strides[i], anchor_stride)
                           for i in range(len(scales))], axis=0)


def norm_boxes(boxes, shape):
    h, w = shape
    return ((boxes - np.array([0, 0, 1, 1])) / np.array([h - 1, w - 1, h - 1, w - 1])).astype(np.float32)


def denorm_boxes(boxes, shape):
    h, w = shape
    return np.around(np.multiply(boxes, np.array([h - 1, w - 1, h - 1, w - 1])) + np.array([0, 0, 1, 1])).astype(np.int32)


# --------------------------------------------------------------------------------------------
# tiles (reference: mrcnn/utils.py:1254-1328)
# --------------------------------------------------------------------------------------------

def generate_tiles(img_xmin, img_xmax, img_ymin, img_ymax, tileSizeX, tileSizeY, gridStepSizeX, gridStepSizeY):
    """Tile coordinates (xmin, xmax, ymin, ymax) with exclusive maxima, row-major over the image range; None when the
    arguments are invalid. (The reference's error branches call an undefined `logger` and raise NameError; the
    intended `return None` is what is implemented here.)"""
    if img_xmax <= img_xmin:
        logger.error("xmax must be > xmin!")
        return None
    if img_ymax <= img_ymin:
        logger.error("ymax must be > ymin!")
        return None
    if tileSizeX <= 0 or tileSizeY <= 0:
        logger.error("Invalid box size given!")
        return None
    if gridStepSizeX <= 0 or gridStepSizeY <= 0 or gridStepSizeX > 1 or gridStepSizeY > 1:
        logger.error("Invalid grid step size given (null or negative)!")
        return None
    Nx = img_xmax - img_xmin + 1
    Ny = img_ymax - img_ymin + 1
    if tileSizeX > Nx or tileSizeY > Ny:
        logger.warning("Invalid box size given (too small or larger than image size)!")
        return None
    stepSizeX = int(np.round(gridStepSizeX * tileSizeX))
    stepSizeY = int(np.round(gridStepSizeY * tileSizeY))
    if stepSizeX < 1 or stepSizeY < 1:          # the reference would loop forever here
        logger.error("Grid step rounds to zero pixels!")
        return None
    spans = []
    for n, size, step in ((Nx, tileSizeX, stepSizeX), (Ny, tileSizeY, stepSizeY)):
        lo_hi, index = [], 0
        while index < n:                      # a tile starts at every step until the range is exhausted
            lo_hi.append((index, index + min(size, n - index)))
            index += step
        spans.append(lo_hi)
    return [(img_xmin + x0, img_xmin + x1, img_ymin + y0, img_ymin + y1) for (y0, y1) in spans[1] for (x0, x1) in spans[0]]

