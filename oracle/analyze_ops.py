"""CPU oracle of the Analyzer mask post-processing. TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and the cpu_baseline leg of the benchmarks may import this module; the
product (caesar-mrcnn_b200/mrcnn/analyze.py + csrc/analyze.cu) never does.

Restates, with full-frame numpy arrays exactly as the reference holds them:
  Analyzer.extract_det_masks                 /root/reference/mrcnn/analyze.py:1162-1423
  Analyzer.make_json_results                 analyze.py:1866-1942   (all keys except "vertexes", see below)
  Analyzer.merge_masks                       analyze.py:2142-2146
  Analyzer.extract_mask_connected_components analyze.py:2148-2151
  Analyzer.are_mask_connected                analyze.py:2154-2173
  Graph (DFS connected components)           /root/reference/mrcnn/graph.py:5-45
  utils.extract_bboxes                       /root/reference/mrcnn/utils.py:33-59

Third-party pieces the reference calls and that are NOT vendored under /root/reference:
  * skimage.measure.label(mask, background=0, return_num=True, connectivity=1) — scikit-image is absent from this
    image; restated with scipy.ndimage.label (default cross-shaped structuring element = 4-connectivity). Both
    number components in raster order of their first pixel.
  * sklearn.metrics.jaccard_score(a.flatten(), b.flatten(), average='binary') — scikit-learn IS installed; the
    oracle calls it when importable and otherwise uses the published formula tp / (tp + fp + fn) in float64
    (0.0 when the denominator is 0); tests check that the two agree.
  * networkx.find_cliques — networkx is installed and called directly (the reference does the same).
  * skimage.measure.find_contours (the "vertexes" key of the JSON objects) — absent and not restated: PARITY
    UNPINNED for that one key; `make_json_results(..., find_contours=None)` emits an empty list for it.

Pinned against the reference itself: tests/golden/make_golden_analyzer.py imports the reference's Analyzer in the
build container (with skimage.measure.label replaced by the scipy call above and find_contours by a stub) and
stores its inputs/outputs in tests/golden/analyzer_golden.json; tests/test_oracle_analyzer.py replays them here.
Note that the reference's score arithmetic (`score_avg += score; score_avg *= 1./n`) follows numpy's scalar
promotion rules: the goldens were generated with the numpy of this image (2.x, NEP 50: float32 stays float32).
"""
import numpy as np

GALAXY_LABELS = ("galaxy_C2", "galaxy_C3", "galaxy", "extended-multisland")   # analyze.py:1223


def label_components(mask):
    """analyze.py:2148-2151 (skimage.measure.label, connectivity=1) -> (labels, ncomponents)."""
    from scipy import ndimage
    labels, n = ndimage.label(np.asarray(mask) != 0)
    return labels, int(n)


def merge_masks(mask1, mask2):
    """analyze.py:2142-2146."""
    mask = mask1 + mask2
    mask[mask > 1] = 1
    return mask


def are_mask_connected(mask1, mask2):
    """analyze.py:2154-2173: connected unless the sum has exactly ncomp1 + ncomp2 components."""
    _, n1 = label_components(mask1)
    _, n2 = label_components(mask2)
    _, n = label_components(merge_masks(mask1, mask2))
    return n != n1 + n2


def jaccard_formula(mask1, mask2):
    a = np.asarray(mask1).ravel() != 0
    b = np.asarray(mask2).ravel() != 0
    tp = np.count_nonzero(a & b)
    den = np.count_nonzero(a | b)
    return np.float64(tp) / np.float64(den) if den else 0.0


def jaccard_binary(mask1, mask2):
    """analyze.py:1272 / 1343: sklearn jaccard_score(average='binary') on the flattened masks."""
    try:
        from sklearn.metrics import jaccard_score
    except ImportError:
        return jaccard_formula(mask1, mask2)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return jaccard_score(np.asarray(mask1).flatten(), np.asarray(mask2).flatten(), average="binary")


class Graph:
    """graph.py:5-45 — undirected graph, recursive pre-order DFS in adjacency insertion order."""

    def __init__(self, n):
        self.n = n
        self.adj = [[] for _ in range(n)]

    def add_edge(self, v, w):
        self.adj[v].append(w)
        self.adj[w].append(v)

    def connected_components(self):
        seen = [False] * self.n
        out = []

        def visit(v, acc):
            seen[v] = True
            acc.append(v)
            for u in self.adj[v]:
                if not seen[u]:
                    visit(u, acc)

        for v in range(self.n):
            if not seen[v]:
                acc = []
                visit(v, acc)
                out.append(acc)
        return out


def extract_bbox(mask):
    """utils.py:33-59 for one mask -> [y1, x1, y2, x2] int32 (y2/x2 exclusive; zeros when empty)."""
    m = np.asarray(mask) != 0
    cols = np.where(np.any(m, axis=0))[0]
    rows = np.where(np.any(m, axis=1))[0]
    if cols.shape[0]:
        x1, x2 = cols[[0, -1]]
        y1, y2 = rows[[0, -1]]
        x2 += 1
        y2 += 1
    else:
        x1, x2, y1, y2 = 0, 0, 0, 0
    return np.array([y1, x1, y2, x2], dtype=np.int32)


def extract_det_masks(masks, n_boxes, class_ids, scores, class_names, score_thr=0.7, split_masks=False,
                      merge_overlapped_masks=True, select_best_overlapped_masks=True, split_source_sidelobe=True,
                      merge_overlap_iou_thr=0.3):
    """analyze.py:1162-1423. masks [H,W,N]; returns dict(masks_final, class_ids_final, class_names_final,
    scores_final, bboxes, captions) — the attributes the reference sets."""
    import networkx as nx

    # score filter, analyze.py:1181-1198
    masks_sel, class_ids_sel, scores_sel = [], [], []
    for i in range(n_boxes):
        if scores[i] < score_thr:
            continue
        masks_sel.append(masks[:, :, i])
        class_ids_sel.append(class_ids[i])
        scores_sel.append(scores[i])
    order = np.argsort(scores_sel)[::-1]                  # :1203

    # optional split into connected components, :1211-1255
    masks_det, class_ids_det, scores_det = [], [], []
    for index in order:
        mask, class_id, score = masks_sel[index], class_ids_sel[index], scores_sel[index]
        if not split_masks or class_names[class_id] in GALAXY_LABELS:
            masks_det.append(mask)
            class_ids_det.append(class_id)
            scores_det.append(score)
            continue
        labels, ncomp = label_components(mask)
        for c in range(ncomp):
            masks_det.append(np.where(labels == c + 1, [1], [0]))
            class_ids_det.append(class_id)
            scores_det.append(score)

    # merge connected same-class masks above the IOU threshold, :1258-1320
    masks_merged, class_ids_merged, scores_merged = [], [], []
    if merge_overlapped_masks:
        n = len(masks_det)
        g = Graph(n)
        for i in range(n):
            for j in range(i + 1, n):
                connected = are_mask_connected(masks_det[i], masks_det[j])
                same_class = class_ids_det[i] == class_ids_det[j]
                iou = jaccard_binary(masks_det[i], masks_det[j])
                if connected and same_class and iou >= merge_overlap_iou_thr:
                    g.add_edge(i, j)
        for comp in g.connected_components():
            if not comp:
                continue
            score_avg = 0
            for j, index in enumerate(comp):
                class_id = class_ids_det[index]
                score_avg += scores_det[index]
                merged = masks_det[index] if j == 0 else merge_masks(merged, masks_det[index])
            score_avg *= 1. / len(comp)
            masks_merged.append(merged)
            class_ids_merged.append(class_id)          # the class of the LAST member (they are all equal anyway)
            scores_merged.append(score_avg)
    else:
        masks_merged, class_ids_merged, scores_merged = list(masks_det), list(class_ids_det), list(scores_det)

    out = dict(masks_final=[], class_ids_final=[], class_names_final=[], scores_final=[], bboxes=[], captions=[])
    if not select_best_overlapped_masks:                 # :1324 — nothing is published otherwise
        return out

    # best of overlapping objects via maximal cliques, :1328-1395
    n = len(masks_merged)
    g_final = nx.Graph()
    for i in range(n):
        label_i = class_names[class_ids_merged[i]]
        for j in range(i + 1, n):
            label_j = class_names[class_ids_merged[j]]
            connected = are_mask_connected(masks_merged[i], masks_merged[j])
            sidelobe_other = (label_i == "spurious") != (label_j == "spurious")
            mergeable = connected
            if connected and split_source_sidelobe and sidelobe_other:
                if jaccard_binary(masks_merged[i], masks_merged[j]) < merge_overlap_iou_thr:
                    mergeable = False
            if mergeable:
                g_final.add_edge(i, j)
    cliques = list(nx.find_cliques(g_final))
    best_score, best_index = [], []
    for clique in cliques:
        top, top_index = -1, -1
        for index in clique:
            if scores_merged[index] > top:
                top, top_index = scores_merged[index], index
        best_score.append(top)
        best_index.append(top_index)
    selected = [True] * n
    for k in sorted(range(len(cliques)), key=lambda q: best_score[q], reverse=True):
        for index in cliques[k]:
            if index != best_index[k] and selected[index]:
                selected[index] = False

    # bounding boxes and publication, :1398-1421
    for index in range(n):
        if not selected[index]:
            continue
        bbox = extract_bbox(masks_merged[index])
        if bbox[1] >= bbox[3] or bbox[0] >= bbox[2]:
            continue
        label = class_names[class_ids_merged[index]]
        out["masks_final"].append(masks_merged[index])
        out["class_ids_final"].append(class_ids_merged[index])
        out["class_names_final"].append(label)
        out["scores_final"].append(scores_merged[index])
        out["bboxes"].append(bbox)
        out["captions"].append("{} {:.2f}".format(label, scores_merged[index]))
    return out


def make_json_results(det, class_names, image_shape, image_id=-1, xmin=0, ymin=0, obj_name_tag="", find_contours=None):
    """analyze.py:1866-1942. `det` = output of extract_det_masks; image_shape = image.shape."""
    results = {"image_id": image_id, "objs": []}
    ny, nx_ = image_shape[0], image_shape[1]
    for i, mask in enumerate(det["masks_final"]):
        class_id = int(det["class_ids_final"][i])
        y1, x1, y2, x2 = (int(v) for v in det["bboxes"][i])
        at_edge = (x1 <= 0 or x1 >= nx_ - 1 or x2 <= 0 or x2 >= nx_ - 1 or
                   y1 <= 0 or y1 >= ny - 1 or y2 <= 0 or y2 >= ny - 1)
        pixels = np.argwhere(mask == 1).tolist()
        if xmin != 0 or ymin != 0:
            for p in pixels:
                p[0] += ymin
                p[1] += xmin
        vertexes = []
        if find_contours is not None:
            padded = np.zeros((mask.shape[0] + 2, mask.shape[1] + 2), dtype=np.uint8)
            padded[1:-1, 1:-1] = mask
            for verts in find_contours(padded, 0.5):
                vertexes.append((np.fliplr(verts) - 1).tolist())
            if xmin != 0 or ymin != 0:
                for contour in vertexes:
                    for v in contour:
                        v[0] += xmin
                        v[1] += ymin
        results["objs"].append({
            "name": "S" + str(i + 1) + "_" + obj_name_tag,
            "x1": xmin + x1, "x2": xmin + x2, "y1": ymin + y1, "y2": ymin + y2,
            "class_id": class_id, "class_name": class_names[class_id], "score": det["scores_final"][i],
            "pixels": pixels, "vertexes": vertexes, "edge": at_edge,
        })
    return results
