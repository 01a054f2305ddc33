"""Shared helpers for the Analyzer parity tests: golden-case decoding and result summarising."""
import hashlib
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "analyzer_golden.json")


def load_golden():
    with open(GOLDEN) as f:
        return json.load(f)


def draw_shapes(shape, shapes):
    m = np.zeros(shape, dtype=bool)
    yy, xx = np.mgrid[0:shape[0], 0:shape[1]]
    for s in shapes:
        if s[0] == "rect":
            m[s[1]:s[3], s[2]:s[4]] = True
        else:
            m |= (yy - s[1]) ** 2 + (xx - s[2]) ** 2 <= s[3] ** 2
    return m


def case_inputs(case):
    H, W, n = case["H"], case["W"], len(case["dets"])
    masks = np.zeros((H, W, n), dtype=bool)
    for i, shapes in enumerate(case["dets"]):
        masks[:, :, i] = draw_shapes((H, W), shapes)
    scores = np.array([float.fromhex(h) for h in case["scores_hex"]], dtype=np.float32)
    class_ids = np.asarray(case["class_ids"], dtype=np.int32)
    return masks, class_ids, scores


def summarise(objs, masks_final, captions, with_pixels):
    """JSON objs (+ final masks / captions) -> the record layout stored in the golden file."""
    out = []
    for i, obj in enumerate(objs):
        px = np.asarray(obj["pixels"], dtype=np.int32).reshape(-1, 2)
        rec = {k: obj[k] for k in ("name", "class_id", "class_name", "edge")}
        for k in ("x1", "x2", "y1", "y2"):
            rec[k] = int(obj[k])
        rec["score_hex"] = float(obj["score"]).hex()
        rec["score_type"] = type(obj["score"]).__name__
        rec["npix"] = int(px.shape[0])
        rec["pixels_sha1"] = hashlib.sha1(np.ascontiguousarray(px).tobytes()).hexdigest()
        if with_pixels:
            rec["pixels"] = px.tolist()
        rec["mask_dtype"] = str(np.asarray(masks_final[i]).dtype)
        rec["caption"] = captions[i]
        out.append(rec)
    return out


def random_detections(rng, H, W, n, n_classes=6, density=1.0):
    """Blobby random masks (unions of rectangles / discs / noise) for oracle-vs-GPU comparisons."""
    masks = np.zeros((H, W, n), dtype=bool)
    yy, xx = np.mgrid[0:H, 0:W]
    for i in range(n):
        for _ in range(int(rng.integers(1, 4))):
            kind = rng.random()
            if kind < 0.4:
                y1, x1 = int(rng.integers(0, H - 1)), int(rng.integers(0, W - 1))
                masks[y1:y1 + int(rng.integers(1, max(2, int(H * 0.3 * density)))),
                      x1:x1 + int(rng.integers(1, max(2, int(W * 0.3 * density)))), i] = True
            elif kind < 0.8:
                cy, cx, r = int(rng.integers(0, H)), int(rng.integers(0, W)), int(rng.integers(1, max(2, int(H * 0.15 * density))))
                masks[:, :, i] |= (yy - cy) ** 2 + (xx - cx) ** 2 <= r * r
            else:
                y1, x1 = int(rng.integers(0, max(1, H - 8))), int(rng.integers(0, max(1, W - 8)))
                masks[y1:y1 + 8, x1:x1 + 8, i] |= rng.random((min(8, H - y1), min(8, W - x1))) < 0.5
    class_ids = rng.integers(1, n_classes, size=n).astype(np.int32)
    scores = rng.uniform(0.5, 1.0, size=n).astype(np.float32)
    return masks, class_ids, scores
