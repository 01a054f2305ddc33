"""End-to-end detection parity metrics (numpy only; neutral helper shared by tests/, bench.py's
cpu_baseline leg and tools/): how far are the final detections of one run from those of another run
of the same graph on the same inputs and weights (here: the bf16 sm_100a engine against the fp32 CPU
oracle, reference outputs mrcnn/model.py:2156-2158 + unmold_detections :2558-2621)?

A detection of run A is *matched* to the not-yet-used detection of run B of the same class with the
largest box IoU, if that IoU is >= iou_match.  Reported per matched pair: |d box| in molded-image
pixels (float boxes of the `detections` tensor times (S-1)), |d score|, IoU of the full-frame masks.
"""
import numpy as np


def box_iou(a, b):
    """a [4], b [M,4] (y1,x1,y2,x2) -> IoU [M]"""
    y1 = np.maximum(a[0], b[:, 0])
    x1 = np.maximum(a[1], b[:, 1])
    y2 = np.minimum(a[2], b[:, 2])
    x2 = np.minimum(a[3], b[:, 3])
    inter = np.maximum(y2 - y1, 0) * np.maximum(x2 - x1, 0)
    ua = (a[2] - a[0]) * (a[3] - a[1]) + (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1]) - inter
    return np.where(ua > 0, inter / np.where(ua > 0, ua, 1), 0.0)


def match_image(boxes_a, cls_a, score_a, boxes_b, cls_b, score_b, masks_a=None, masks_b=None, px_scale=1.0,
                iou_match=0.5):
    """boxes_* [N,4] (y1,x1,y2,x2), cls_* [N], score_* [N] of the real detections of two runs (either the non-padded
    rows of the `detections` graph output, px_scale = S-1, or the unmolded integer boxes, px_scale = 1);
    masks_* optional [H,W,N] aligned with the same rows.
    Returns dict(n_a, n_b, pairs=[(ia, ib, dbox_px, dscore, box_iou, mask_iou|None)])."""
    a = np.asarray(boxes_a, np.float64).reshape(-1, 4)
    b = np.asarray(boxes_b, np.float64).reshape(-1, 4)
    cls_a, cls_b = np.asarray(cls_a), np.asarray(cls_b)
    score_a, score_b = np.asarray(score_a, np.float64), np.asarray(score_b, np.float64)
    used = np.zeros(len(b), bool)
    pairs = []
    for ia in np.argsort(-score_a, kind="stable"):
        if not len(b):
            break
        iou = box_iou(a[ia], b)
        iou = np.where((cls_b == cls_a[ia]) & ~used, iou, -1.0)
        ib = int(np.argmax(iou))
        if iou[ib] < iou_match:
            continue
        used[ib] = True
        dbox = float(np.abs(a[ia] - b[ib]).max() * px_scale)
        dscore = float(abs(score_a[ia] - score_b[ib]))
        miou = None
        if masks_a is not None and masks_b is not None:
            ma, mb = np.asarray(masks_a[:, :, ia], bool), np.asarray(masks_b[:, :, ib], bool)
            un = np.logical_or(ma, mb).sum()
            miou = float(np.logical_and(ma, mb).sum() / un) if un else 1.0
        pairs.append((int(ia), ib, dbox, dscore, float(iou[ib]), miou))
    return {"n_a": int(len(a)), "n_b": int(len(b)), "pairs": pairs}


def match_detections_tensor(det_a, det_b, size):
    """`detections` graph outputs [D,6] (normalised boxes, class, score; zero padded) of two runs"""
    a = det_a[det_a[:, 4] > 0]
    b = det_b[det_b[:, 4] > 0]
    return match_image(a[:, :4], a[:, 4], a[:, 5], b[:, :4], b[:, 4], b[:, 5], px_scale=size - 1)


def match_results(res_a, res_b):
    """detect()-style dicts (rois int32 [N,4], class_ids, scores, masks [H,W,N]) of two runs"""
    return match_image(res_a["rois"], res_a["class_ids"], res_a["scores"], res_b["rois"], res_b["class_ids"],
                       res_b["scores"], res_a["masks"], res_b["masks"], px_scale=1.0)


def summarize(per_image):
    """list of match_image results -> one flat dict of distribution statistics."""
    n_a = sum(r["n_a"] for r in per_image)
    n_b = sum(r["n_b"] for r in per_image)
    pairs = [p for r in per_image for p in r["pairs"]]
    out = {"images": len(per_image), "detections_a": n_a, "detections_b": n_b, "matched": len(pairs),
           "matched_frac": len(pairs) / max(1, max(n_a, n_b))}
    if pairs:
        dbox = np.array([p[2] for p in pairs])
        dsc = np.array([p[3] for p in pairs])
        out.update({"dbox_px_median": float(np.median(dbox)), "dbox_px_p95": float(np.percentile(dbox, 95)),
                    "dbox_px_max": float(dbox.max()), "dscore_median": float(np.median(dsc)),
                    "dscore_p95": float(np.percentile(dsc, 95)), "dscore_max": float(dsc.max()),
                    "same_rank_frac": float(np.mean([p[0] == p[1] for p in pairs]))})
        mi = np.array([p[5] for p in pairs if p[5] is not None])
        if len(mi):
            out.update({"mask_iou_median": float(np.median(mi)), "mask_iou_p05": float(np.percentile(mi, 5)),
                        "mask_iou_min": float(mi.min()), "mask_iou_ge_0.99_frac": float(np.mean(mi >= 0.99)),
                        "mask_iou_ge_0.9_frac": float(np.mean(mi >= 0.9))})
    return out


def stage_errors(out_a, out_b):
    """max / rms absolute differences of the float graph outputs of two runs (dicts keyed like
    mrcnn/model.py:2156-2158)."""
    res = {}
    for k in ("rpn_class", "rpn_bbox", "mrcnn_class", "mrcnn_bbox"):
        if k in out_a and k in out_b:
            d = np.abs(np.asarray(out_a[k], np.float64) - np.asarray(out_b[k], np.float64))
            res[k + "_max_abs"] = float(d.max())
            res[k + "_rms"] = float(np.sqrt((d ** 2).mean()))
    return res
