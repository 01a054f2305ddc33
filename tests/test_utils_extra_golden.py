"""CPU: the host box / metric helpers of mrcnn.utils (compute_iou, compute_overlaps_masks, non_max_suppression,
apply_box_deltas, compute_matches, compute_ap, compute_ap_range, compute_recall) against outputs of the reference's OWN
functions (tests/golden/ref_utils_extra_golden.npz, written by tests/golden/make_golden_utils_extra.py from
/root/reference/mrcnn/utils.py:75-863) — bit for bit, dtypes included."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_utils_extra_golden.npz")


@pytest.fixture(scope="module")
def g():
    return np.load(GOLDEN)


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


def test_box_helpers_match_the_reference(g):
    from mrcnn import utils
    for k in range(4):
        boxes, scores, deltas = g["nms%d_boxes" % k], g["nms%d_scores" % k], g["nms%d_deltas" % k]
        area = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])
        assert same(utils.compute_iou(boxes[0], boxes, area[0], area), g["nms%d_iou0" % k])
        for thr in (0.3, 0.7):
            keep = utils.non_max_suppression(boxes, scores, thr)
            assert same(keep, g["nms%d_keep_%d" % (k, int(thr * 10))]) and keep.dtype == np.int32
        assert same(utils.apply_box_deltas(boxes, deltas), g["nms%d_applied" % k])
    got = np.array([utils.get_iou(list(p[:4]), list(p[4:])) for p in g["get_iou_pairs"]], dtype=np.float64)
    assert same(got, g["get_iou_values"]) and (got == 0).any() and (got > 0.05).any()
    with pytest.raises(AssertionError):
        utils.get_iou([0, 0, 0, 5], [0, 0, 5, 5])
    with pytest.raises(AssertionError):
        utils.non_max_suppression(np.zeros((0, 4), np.float32), np.zeros(0, np.float32), 0.5)


def test_mask_metrics_match_the_reference(g):
    from mrcnn import utils
    done = 0
    for k in range(5):
        a = {n: g["ap%d_%s" % (k, n)] for n in ("gt_boxes", "gt_cls", "gt_masks", "p_boxes", "p_cls", "p_scores", "p_masks")}
        assert same(utils.compute_overlaps_masks(a["p_masks"], a["gt_masks"]), g["ap%d_overlaps_masks" % k])
        if "ap%d_gt_match" % k not in g.files:
            continue
        args = (a["gt_boxes"], a["gt_cls"], a["gt_masks"], a["p_boxes"], a["p_cls"], a["p_scores"], a["p_masks"])
        gm, pm, ov = utils.compute_matches(*args, 0.5, 0.0)
        assert same(gm, g["ap%d_gt_match" % k]) and same(pm, g["ap%d_pred_match" % k]) and same(ov, g["ap%d_overlaps" % k])
        gm, pm, _ = utils.compute_matches(*args, 0.3, 0.4)
        assert same(gm, g["ap%d_gt_match_b" % k]) and same(pm, g["ap%d_pred_match_b" % k])
        ap, prec, rec, _ = utils.compute_ap(*args, 0.5)
        assert same(np.array([ap]), g["ap%d_ap" % k]) and same(prec, g["ap%d_prec" % k]) and same(rec, g["ap%d_rec" % k])
        assert same(np.array([utils.compute_ap_range(*args, verbose=0)]), g["ap%d_ap_range" % k])
        rc, pos = utils.compute_recall(a["p_boxes"], utils.trim_zeros(a["gt_boxes"]), 0.5)
        assert same(np.array([rc]), g["ap%d_recall" % k]) and same(pos, g["ap%d_recall_ids" % k])
        done += 1
    assert done >= 3
