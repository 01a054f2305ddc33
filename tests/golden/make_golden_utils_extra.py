#!/usr/bin/env python
"""Generates tests/golden/ref_utils_extra_golden.npz by running the REAL reference functions of mrcnn/utils.py that are pure
numpy (compute_iou, compute_overlaps_masks, non_max_suppression, apply_box_deltas, compute_matches, compute_ap,
compute_ap_range, compute_recall) on seeded inputs.  Build container only (needs /root/reference; the absent packages are
stubbed exactly as in make_golden_from_reference.py).  Nothing here is imported by the product."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_from_reference as G  # noqa: E402

OUT = os.path.join(HERE, "ref_utils_extra_golden.npz")


def instances(rng, H, W, n, n_classes=4):
    boxes = np.zeros((n, 4), np.int32)
    masks = np.zeros((H, W, n), bool)
    for i in range(n):
        h, w = rng.integers(3, H // 2), rng.integers(3, W // 2)
        y, x = rng.integers(0, H - h), rng.integers(0, W - w)
        boxes[i] = [y, x, y + h, x + w]
        yy, xx = np.mgrid[0:H, 0:W]
        masks[:, :, i] = ((yy - (y + h / 2)) / (h / 2)) ** 2 + ((xx - (x + w / 2)) / (w / 2)) ** 2 <= 1
    return boxes, rng.integers(1, n_classes, size=n).astype(np.int32), masks


def main():
    G.install_stubs()
    sys.path.insert(0, G.REF)
    import logging
    logging.disable(logging.CRITICAL)
    from mrcnn import utils
    rng = np.random.default_rng(2024)
    g = {}
    # ---- compute_iou / non_max_suppression / apply_box_deltas
    for k, n in enumerate((1, 7, 60, 300)):
        y1, x1 = rng.uniform(0, 200, n), rng.uniform(0, 200, n)
        boxes = np.stack([y1, x1, y1 + rng.uniform(1, 60, n), x1 + rng.uniform(1, 60, n)], axis=1).astype(np.float32)
        if k == 2:
            boxes = np.round(boxes).astype(np.int32)               # integer boxes: the float32 conversion branch
        scores = rng.uniform(0, 1, n).astype(np.float32)
        if n >= 60:
            scores[::5] = scores[0]                                 # ties
        deltas = (rng.normal(size=(n, 4)) * 0.2).astype(np.float32)
        area = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])
        g["nms%d_boxes" % k], g["nms%d_scores" % k], g["nms%d_deltas" % k] = boxes, scores, deltas
        g["nms%d_iou0" % k] = utils.compute_iou(boxes[0], boxes, area[0], area)
        for thr in (0.3, 0.7):
            g["nms%d_keep_%d" % (k, int(thr * 10))] = utils.non_max_suppression(boxes, scores, thr)
        g["nms%d_applied" % k] = utils.apply_box_deltas(boxes, deltas)
    # ---- get_iou (pairs of integer and float boxes, disjoint / touching / nested)
    pairs = []
    for _ in range(40):
        a = rng.integers(0, 50, size=2)
        b = rng.integers(0, 50, size=2)
        bb1 = [int(a[0]), int(a[1]), int(a[0] + rng.integers(1, 30)), int(a[1] + rng.integers(1, 30))]
        bb2 = [int(b[0]), int(b[1]), int(b[0] + rng.integers(1, 30)), int(b[1] + rng.integers(1, 30))]
        pairs.append(bb1 + bb2)
    pairs.append([0, 0, 10, 10, 10, 10, 20, 20])
    pairs.append([0, 0, 10, 10, 2, 2, 5, 5])
    pairs = np.array(pairs, dtype=np.float64)
    pairs[::3] += 0.25
    g["get_iou_pairs"] = pairs
    g["get_iou_values"] = np.array([utils.get_iou(list(p[:4]), list(p[4:])) for p in pairs], dtype=np.float64)
    # ---- mask overlaps / matches / AP / recall
    for k, (n_gt, n_pred) in enumerate(((3, 5), (6, 6), (4, 0), (0, 3), (8, 12))):
        H = W = 48
        gt_boxes, gt_cls, gt_masks = instances(rng, H, W, n_gt)
        p_boxes, p_cls, p_masks = instances(rng, H, W, n_pred)
        m = min(n_gt, n_pred)
        if m:                                                        # some predictions are jittered copies of GT instances
            p_boxes[:m], p_cls[:m] = gt_boxes[:m], gt_cls[:m]
            p_masks[:, :, :m] = np.roll(gt_masks[:, :, :m], 1, axis=0)
        p_scores = rng.uniform(0.1, 1, n_pred).astype(np.float32)
        if k == 4:                                                   # zero padding at the end, as detect() batches carry it
            gt_boxes = np.concatenate([gt_boxes, np.zeros((2, 4), np.int32)])
            gt_cls = np.concatenate([gt_cls, np.zeros(2, np.int32)])
            gt_masks = np.concatenate([gt_masks, np.zeros((H, W, 2), bool)], axis=-1)
        for name, val in (("gt_boxes", gt_boxes), ("gt_cls", gt_cls), ("gt_masks", gt_masks), ("p_boxes", p_boxes), ("p_cls", p_cls),
                          ("p_scores", p_scores), ("p_masks", p_masks)):
            g["ap%d_%s" % (k, name)] = val
        g["ap%d_overlaps_masks" % k] = utils.compute_overlaps_masks(p_masks, gt_masks)
        if n_gt and n_pred:
            gm, pm, ov = utils.compute_matches(gt_boxes, gt_cls, gt_masks, p_boxes, p_cls, p_scores, p_masks, 0.5, 0.0)
            g["ap%d_gt_match" % k], g["ap%d_pred_match" % k], g["ap%d_overlaps" % k] = gm, pm, ov
            gm2, pm2, _ = utils.compute_matches(gt_boxes, gt_cls, gt_masks, p_boxes, p_cls, p_scores, p_masks, 0.3, 0.4)
            g["ap%d_gt_match_b" % k], g["ap%d_pred_match_b" % k] = gm2, pm2
            ap, prec, rec, _ = utils.compute_ap(gt_boxes, gt_cls, gt_masks, p_boxes, p_cls, p_scores, p_masks, 0.5)
            g["ap%d_ap" % k], g["ap%d_prec" % k], g["ap%d_rec" % k] = np.array([ap]), prec, rec
            g["ap%d_ap_range" % k] = np.array([utils.compute_ap_range(gt_boxes, gt_cls, gt_masks, p_boxes, p_cls, p_scores, p_masks,
                                                                       verbose=0)])
            rc, pos = utils.compute_recall(p_boxes, utils.trim_zeros(gt_boxes), 0.5)
            g["ap%d_recall" % k], g["ap%d_recall_ids" % k] = np.array([rc]), pos
    np.savez_compressed(OUT, **g)
    print("wrote", OUT, len(g), "arrays", os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
