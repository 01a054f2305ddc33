"""CPU: the Analyzer oracle (oracle/analyze_ops.py) against outputs of the REAL reference Analyzer
(tests/golden/make_golden_analyzer.py ran /root/reference/mrcnn/analyze.py in the build container), plus the
equivalences the CUDA path relies on."""
import numpy as np
import pytest

from oracle import analyze_ops as A
import analyzer_cases as C


@pytest.fixture(scope="module")
def golden_cases():
    return C.load_golden()


def test_oracle_matches_reference_goldens(golden_cases):
    names = golden_cases["class_names"]
    assert len(golden_cases["cases"]) >= 50
    for case in golden_cases["cases"]:
        masks, class_ids, scores = C.case_inputs(case)
        det = A.extract_det_masks(masks, masks.shape[2], class_ids, scores, names, **case["options"])
        xmin, ymin = case["origin"]
        res = A.make_json_results(det, names, (case["H"], case["W"], 3), image_id=case["name"], xmin=xmin, ymin=ymin,
                                  obj_name_tag="t0")
        got = C.summarise(res["objs"], det["masks_final"], det["captions"], case["H"] * case["W"] <= 64 * 64)
        assert got == case["objs"], (case["name"], case["options"], case["origin"])


def test_touch_test_is_equivalent_to_component_counting():
    """are_mask_connected (three labellings) == 'a pixel of one coincides with or is 4-adjacent to a pixel of the other'."""
    rng = np.random.default_rng(5)
    for _ in range(300):
        H, W = int(rng.integers(3, 20)), int(rng.integers(3, 40))
        a = rng.random((H, W)) < rng.uniform(0.02, 0.4)
        b = rng.random((H, W)) < rng.uniform(0.02, 0.4)
        grown = b.copy()
        grown[1:] |= b[:-1]
        grown[:-1] |= b[1:]
        grown[:, 1:] |= b[:, :-1]
        grown[:, :-1] |= b[:, 1:]
        assert A.are_mask_connected(a, b) == bool(np.any(a & grown))


def test_jaccard_formula_matches_sklearn():
    pytest.importorskip("sklearn")
    rng = np.random.default_rng(6)
    for k in range(50):
        a = rng.random((17, 23)) < 0.3
        b = rng.random((17, 23)) < (0.3 if k else 0.0)
        if k == 1:
            a[:] = False
        assert float(A.jaccard_binary(a, b)) == float(A.jaccard_formula(a, b))
        assert float(A.jaccard_binary(a.astype(np.int64), b)) == float(A.jaccard_formula(a, b))


def test_graph_components_are_dfs_preorder():
    g = A.Graph(6)
    for v, w in ((0, 4), (4, 2), (0, 5), (1, 3)):
        g.add_edge(v, w)
    assert g.connected_components() == [[0, 4, 2, 5], [1, 3]]


def test_label_numbering_is_raster_order_of_first_pixel():
    m = np.zeros((5, 7), dtype=bool)
    m[0, 5] = True            # first pixel in raster order
    m[1, 0:3] = True          # second
    m[2:5, 2] = True          # joins the second from below
    m[4, 4:7] = True          # third ... but touches nothing above
    m[3, 6] = True            # joins the third, appears earlier in raster order than (4,4)
    labels, n = A.label_components(m)
    assert n == 3
    assert labels[0, 5] == 1 and labels[1, 0] == 2 and labels[4, 2] == 2 and labels[3, 6] == 3 and labels[4, 4] == 3


def test_extract_bbox_empty_and_full():
    assert A.extract_bbox(np.zeros((4, 4), bool)).tolist() == [0, 0, 0, 0]
    assert A.extract_bbox(np.ones((4, 6), bool)).tolist() == [0, 0, 4, 6]
