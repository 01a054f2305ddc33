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


class _Arr:
    """Minimal stand-in for a device tensor (shape + slicing) used by NumpyPlaneOps."""

    def __init__(self, a):
        self.a = a

    @property
    def shape(self):
        return self.a.shape

    def __getitem__(self, k):
        return _Arr(self.a[k])


class NumpyPlaneOps:
    """TEST DOUBLE of mrcnn.analyze.MaskPlaneOps: the same method contract computed with numpy on full-frame bool
    arrays, so the host logic of mrcnn/analyze.py (ordering, graphs, cliques, selection, result assembly) can be
    exercised by the CPU suite. Never used by the product."""

    class torch:                       # only what analyze_frames touches for its timing marks
        class cuda:
            @staticmethod
            def synchronize():
                pass

    def __init__(self, masks):
        self.masks = masks             # [F,H,W,D] bool, addressed through plane_of exactly like the device block

    def pack(self, ptr, F, H, W, D, plane_of, n):
        out = np.zeros((n, H, W), bool)
        for flat in np.nonzero(plane_of >= 0)[0]:
            out[plane_of[flat]] = self.masks[flat // D, :, :, flat % D]
        return _Arr(out)

    def area_bbox(self, pl, H, W):
        area = pl.a.reshape(len(pl.a), -1).sum(1).astype(np.int32)
        bbox = np.zeros((len(pl.a), 4), np.int32)
        for i, m in enumerate(pl.a):
            ys, xs = np.nonzero(m.any(1))[0], np.nonzero(m.any(0))[0]
            if len(ys):
                bbox[i] = [ys[0], xs[0], ys[-1] + 1, xs[-1] + 1]
        return _Arr(area), _Arr(bbox)

    def pair_stats(self, pl, H, W, pairs, bbox=None):
        a = pl.a
        grown = a.copy()
        grown[:, 1:] |= a[:, :-1]
        grown[:, :-1] |= a[:, 1:]
        grown[:, :, 1:] |= a[:, :, :-1]
        grown[:, :, :-1] |= a[:, :, 1:]
        inter = np.array([np.count_nonzero(a[i] & a[j]) for i, j in pairs], dtype=np.int32).reshape(-1)
        touch = np.array([np.any(a[i] & grown[j]) for i, j in pairs], dtype=np.int32).reshape(-1)
        return _Arr(inter), _Arr(touch)

    def union(self, pl, H, W, groups):
        if isinstance(groups, tuple):
            members, offsets = groups
            groups = [members[offsets[i]:offsets[i + 1]] for i in range(len(offsets) - 1)]
        return _Arr(np.stack([np.any(pl.a[g], axis=0) for g in groups]) if groups else np.zeros((0, H, W), bool))

    def label(self, pl, H, W):
        """4-connected components, numbered in raster order of their first pixel (= skimage.measure.label(connectivity=1))"""
        from scipy import ndimage
        cross = np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]])
        labels = np.zeros((len(pl.a), H, W), np.int32)
        counts = np.zeros(len(pl.a), np.int32)
        for i, m in enumerate(pl.a):
            labels[i], counts[i] = ndimage.label(m, structure=cross)
        return _Arr(labels), _Arr(counts)

    def select(self, labels, H, W, src, comp):
        src, comp = np.asarray(src, dtype=np.int64), np.asarray(comp, dtype=np.int64)
        return _Arr(np.stack([labels.a[s_] == c for s_, c in zip(src, comp)]) if len(src) else np.zeros((0, H, W), bool))

    def concat(self, a, b):
        return _Arr(np.concatenate([a.a, b.a], axis=0))

    def gather(self, pl, index):
        return _Arr(pl.a[np.asarray(index, dtype=np.int64)])

    def pixels(self, pl, H, W, areas, y0=0, x0=0):
        offsets = np.zeros(len(areas) + 1, np.int64)
        offsets[1:] = np.cumsum(areas)
        px = np.concatenate([np.argwhere(m) for m in pl.a] + [np.zeros((0, 2), np.int64)]).astype(np.int32)
        return px + np.array([y0, x0], dtype=np.int32), offsets

    def unpack(self, pl, H, W):
        return pl.a.astype(np.uint8)

    def host(self, t):
        return t.a
