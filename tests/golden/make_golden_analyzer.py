#!/usr/bin/env python
"""Generates tests/golden/analyzer_golden.json by running the REAL reference Analyzer.

Runs only in the build container (needs /root/reference). The reference's mrcnn/analyze.py is imported with
inert stubs for the packages that are absent here (TensorFlow, Keras, scikit-image, OpenCV, imutils, astropy,
matplotlib, regions); networkx, scikit-learn and numpy are the real ones. Two skimage functions are called by the
code under test:
  * skimage.measure.label  -> replaced by scipy.ndimage.label with the default cross structuring element
    (4-connectivity, raster-order numbering — the semantics of connectivity=1), recorded here;
  * skimage.measure.find_contours -> stub returning no contours, so the "vertexes" key is NOT pinned.
Executed reference code: Analyzer.extract_det_masks (analyze.py:1162-1423), Analyzer.make_json_results
(:1866-1942), merge_masks / extract_mask_connected_components / are_mask_connected (:2142-2173),
mrcnn/graph.py, utils.extract_bboxes (utils.py:33-59).

Every case stores its inputs (masks as per-detection lists of filled rectangles / discs, so the file stays small)
and the reference's outputs (class ids, scores as float.hex, bboxes, per-object pixel checksums and counts, and
the full pixel lists for the small cases). Nothing here is imported by the product.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_from_reference as base  # noqa: E402  (stub machinery)

OUT = os.path.join(HERE, "analyzer_golden.json")
CLASS_NAMES = ["bkg", "spurious", "compact", "extended", "extended-multisland", "flagged"]


def label_stub(mask, background=0, return_num=True, connectivity=1):
    from scipy import ndimage
    assert background == 0 and return_num and connectivity == 1
    labels, n = ndimage.label(np.asarray(mask) != 0)
    return labels, n


def draw_shapes(shape, shapes):
    """shapes: list of ('rect', y1, x1, y2, x2) / ('disc', cy, cx, r) -> bool mask (union)."""
    m = np.zeros(shape, dtype=bool)
    yy, xx = np.mgrid[0:shape[0], 0:shape[1]]
    for s in shapes:
        if s[0] == "rect":
            m[s[1]:s[3], s[2]:s[4]] = True
        else:
            m |= (yy - s[1]) ** 2 + (xx - s[2]) ** 2 <= s[3] ** 2
    return m


def random_case(rng, H, W, n, n_classes=6, tie_scores=False):
    dets = []
    for _ in range(n):
        shapes = []
        for _ in range(int(rng.integers(1, 4))):
            if rng.random() < 0.5:
                y1, x1 = int(rng.integers(0, H - 2)), int(rng.integers(0, W - 2))
                shapes.append(["rect", y1, x1, min(H, y1 + int(rng.integers(1, H // 3))), min(W, x1 + int(rng.integers(1, W // 3)))])
            else:
                shapes.append(["disc", int(rng.integers(0, H)), int(rng.integers(0, W)), int(rng.integers(1, max(2, H // 6)))])
        dets.append(shapes)
    class_ids = rng.integers(1, n_classes, size=n).astype(np.int32)
    scores = rng.uniform(0.5, 1.0, size=n).astype(np.float32)
    if tie_scores:
        scores = np.round(scores * 10).astype(np.float32) / np.float32(10)
    return dets, class_ids, scores


def handmade_cases():
    """Edge cases: empty mask, touching (4-adjacent) but not overlapping, diagonal neighbours (NOT connected),
    spurious-vs-source low IOU, nested masks, image-border objects, ragged width."""
    cases = []
    cases.append(dict(name="touching_and_diagonal", H=20, W=37, dets=[
        [["rect", 2, 2, 6, 6]], [["rect", 2, 6, 6, 10]],            # share an edge: connected, IOU 0
        [["rect", 6, 10, 9, 13]],                                   # diagonal to the second: not connected
        [["rect", 12, 30, 20, 37]],                                 # at the image border, ragged last word
        [],                                                         # empty mask
    ], class_ids=[2, 2, 2, 3, 2], scores=[0.95, 0.9, 0.85, 0.8, 0.99]))
    cases.append(dict(name="merge_same_class_high_iou", H=32, W=64, dets=[
        [["rect", 4, 4, 20, 40]], [["rect", 6, 6, 22, 42]], [["rect", 4, 4, 20, 40]],   # heavy overlap, same class
        [["rect", 5, 5, 21, 41]],                                                     # same place, other class
        [["disc", 26, 55, 3]],
    ], class_ids=[2, 2, 2, 3, 2], scores=[0.91, 0.97, 0.75, 0.93, 0.72]))
    cases.append(dict(name="spurious_vs_source", H=48, W=48, dets=[
        [["rect", 10, 10, 30, 30]], [["rect", 28, 28, 40, 40]],     # small overlap, spurious vs compact: kept apart
        [["rect", 11, 11, 29, 29]],                                 # large overlap with the first, spurious
        [["rect", 0, 0, 3, 48]],
    ], class_ids=[2, 1, 1, 4, ], scores=[0.9, 0.8, 0.95, 0.71]))
    cases.append(dict(name="below_threshold_and_multi_island", H=40, W=70, dets=[
        [["rect", 1, 1, 5, 5], ["rect", 10, 10, 15, 15], ["disc", 30, 50, 6]],   # three islands, compact
        [["rect", 1, 60, 5, 65], ["rect", 20, 60, 25, 65]],                      # two islands, multi-island class
        [["rect", 2, 2, 4, 4]],                                                  # below threshold
        [["rect", 12, 12, 14, 40]],
    ], class_ids=[2, 4, 2, 3], scores=[0.9, 0.88, 0.5, 0.86]))
    return cases


def main():
    base.install_stubs()
    for name in ("cv2", "imutils", "regions", "skimage.draw", "astropy.io.fits", "networkx.nonexistent"):
        if name not in sys.modules and not name.startswith("networkx"):
            base._stub(name)
    sys.modules["skimage.measure"].label = label_stub
    sys.modules["skimage.measure"].find_contours = lambda image, level: []
    sys.modules["skimage"].measure = sys.modules["skimage.measure"]
    sys.modules["skimage"].draw = sys.modules["skimage.draw"]
    sys.path.insert(0, base.REF)
    import logging
    logging.disable(logging.CRITICAL)
    import warnings
    warnings.simplefilter("ignore")
    from mrcnn import analyze as ref_analyze

    class Cfg:
        NUM_CLASSES = len(CLASS_NAMES)
        CLASS_NAMES = CLASS_NAMES

    rng = np.random.default_rng(20261018)
    cases = handmade_cases()
    for k, (H, W, n) in enumerate([(64, 64, 12), (96, 130, 25), (256, 256, 40), (50, 33, 8), (128, 128, 30)]):
        dets, cls, sc = random_case(rng, H, W, n, tie_scores=(k == 3))
        cases.append(dict(name="random_%d" % k, H=H, W=W, dets=dets, class_ids=cls.tolist(), scores=[float(s) for s in sc]))

    option_sets = [
        dict(),                                           # reference defaults: merge + select best
        dict(split_masks=True),
        dict(split_masks=True, merge_overlapped_masks=False),
        dict(split_source_sidelobe=False, merge_overlap_iou_thr=0.1),
        dict(score_thr=0.85),
    ]
    golden = {"numpy": np.__version__, "class_names": CLASS_NAMES, "cases": []}
    for case in cases:
        H, W = case["H"], case["W"]
        n = len(case["dets"])
        masks = np.zeros((H, W, n), dtype=bool)
        for i, shapes in enumerate(case["dets"]):
            masks[:, :, i] = draw_shapes((H, W), [tuple(s) for s in shapes])
        scores = np.asarray(case["scores"], dtype=np.float32)
        class_ids = np.asarray(case["class_ids"], dtype=np.int32)
        for opts in option_sets:
            for origin in ((0, 0), (100, 7)):
                if origin != (0, 0) and opts:
                    continue
                an = ref_analyze.Analyzer(None, Cfg())
                an.class_names = CLASS_NAMES
                an.masks, an.boxes, an.class_ids, an.scores = masks, np.zeros((n, 4), dtype=np.int32), class_ids, scores
                an.nobjects = n
                an.image = np.zeros((H, W, 3), dtype=np.uint8)
                an.image_id = case["name"]
                an.image_xmin, an.image_ymin = origin
                an.obj_name_tag = "t0"
                for key, val in opts.items():
                    setattr(an, key, val)
                an.extract_det_masks()
                an.make_json_results()
                objs = []
                for i, obj in enumerate(an.results["objs"]):
                    px = np.asarray(obj["pixels"], dtype=np.int32).reshape(-1, 2)
                    rec = {k: obj[k] for k in ("name", "x1", "x2", "y1", "y2", "class_id", "class_name", "edge")}
                    rec["x1"], rec["x2"], rec["y1"], rec["y2"] = int(rec["x1"]), int(rec["x2"]), int(rec["y1"]), int(rec["y2"])
                    rec["score_hex"] = float(obj["score"]).hex()
                    rec["score_type"] = type(obj["score"]).__name__
                    rec["npix"] = int(px.shape[0])
                    rec["pixels_sha1"] = hashlib.sha1(np.ascontiguousarray(px).tobytes()).hexdigest()
                    if H * W <= 64 * 64:
                        rec["pixels"] = px.tolist()
                    rec["mask_dtype"] = str(np.asarray(an.masks_final[i]).dtype)
                    rec["caption"] = an.captions[i]
                    objs.append(rec)
                golden["cases"].append(dict(name=case["name"], H=H, W=W, dets=case["dets"], class_ids=case["class_ids"],
                                            scores_hex=[float(s).hex() for s in scores], options=opts,
                                            origin=list(origin), objs=objs))
    with open(OUT, "w") as f:
        json.dump(golden, f, separators=(",", ":"))
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(golden["cases"]), "cases")


if __name__ == "__main__":
    main()
