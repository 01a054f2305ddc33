"""GPU parity of the Analyzer mask post-processing through the C ABI.

Bit-exact integer work: every primitive is compared with the numpy/scipy oracle on seeded masks (ragged widths,
empty and full frames), the whole extract_det_masks + make_json_results pipeline is compared with the outputs of the
REAL reference Analyzer stored in tests/golden/analyzer_golden.json and with the oracle on larger random cases,
and the device-resident batch path is compared with the host-array path."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

import analyzer_cases as C  # noqa: E402
from oracle import analyze_ops as A  # noqa: E402

CLASS_NAMES = ["bkg", "spurious", "compact", "extended", "extended-multisland", "flagged"]


@pytest.fixture(scope="module")
def ops():
    from mrcnn.analyze import MaskPlaneOps
    return MaskPlaneOps(0)


def _planes(ops, masks):
    """masks [H,W,n] bool -> device planes [n, words]"""
    H, W, n = masks.shape
    d = ops.to_dev(masks.view(np.uint8), np.uint8)
    planes = ops.pack(d.data_ptr(), 1, H, W, n, np.arange(n, dtype=np.int32), n)
    torch.cuda.synchronize()
    return planes


def _noise_masks(rng, H, W, n):
    masks = np.zeros((H, W, n), dtype=bool)
    for i in range(n):
        masks[:, :, i] = rng.random((H, W)) < rng.choice([0.0, 0.02, 0.2, 0.5, 0.8, 1.0])
    return masks


@pytest.mark.parametrize("H,W,n", [(5, 7, 3), (20, 37, 9), (33, 64, 12), (64, 65, 20), (256, 256, 100), (130, 300, 7), (50, 100, 16), (17, 40, 8)])
def test_pack_unpack_area_bbox_pixels(ops, H, W, n):
    rng = np.random.default_rng(H * 1000 + W)
    masks = np.concatenate([_noise_masks(rng, H, W, n - 1), np.zeros((H, W, 1), bool)], axis=2)
    planes = _planes(ops, masks)
    assert planes.shape == (n, H * ((W + 31) // 32))
    back = ops.unpack(planes, H, W)
    assert np.array_equal(back.astype(bool), np.moveaxis(masks, 2, 0))
    area, bbox = ops.area_bbox(planes, H, W)
    area, bbox = ops.host(area), ops.host(bbox)
    for i in range(n):
        assert area[i] == masks[:, :, i].sum()
        assert np.array_equal(bbox[i], A.extract_bbox(masks[:, :, i]))
    px, offsets = ops.pixels(planes, H, W, area, 3, 11)
    for i in range(n):
        assert np.array_equal(px[offsets[i]:offsets[i + 1]], np.argwhere(masks[:, :, i] == 1) + np.array([3, 11]))


def test_pack_selects_and_reorders_planes_across_frames(ops):
    rng = np.random.default_rng(3)
    F, H, W, D = 3, 40, 70, 16
    masks = rng.random((F, H, W, D)) < 0.3
    plane_of = np.full(F * D, -1, dtype=np.int32)
    picks = rng.permutation(F * D)[:20]
    plane_of[picks] = np.arange(20)
    d = ops.to_dev(masks.view(np.uint8), np.uint8)
    planes = ops.pack(d.data_ptr(), F, H, W, D, plane_of, 20)
    back = ops.unpack(planes, H, W).astype(bool)
    for m, flat in enumerate(picks):
        assert np.array_equal(back[m], masks[flat // D, :, :, flat % D])


@pytest.mark.parametrize("H,W", [(9, 31), (24, 33), (40, 96), (64, 100)])
def test_pair_stats_and_union(ops, H, W):
    rng = np.random.default_rng(W)
    n = 10
    masks = _noise_masks(rng, H, W, n)
    masks[:, :, 0] = False
    masks[H // 2, :, 1] = True                 # a full row crossing every word boundary
    masks[:, :, 2] = False
    masks[H // 2 - 1, ::2, 2] = True           # touches it from above only
    planes = _planes(ops, masks)
    i, j = np.triu_indices(n, k=1)
    pairs = np.stack([i, j], axis=1).astype(np.int32)
    inter, touch = ops.pair_stats(planes, H, W, pairs)
    inter, touch = ops.host(inter), ops.host(touch)
    _, d_bbox = ops.area_bbox(planes, H, W)                    # bounding-box prefilter: same answers
    inter_b, touch_b = ops.pair_stats(planes, H, W, pairs, d_bbox)
    assert np.array_equal(ops.host(inter_b), inter) and np.array_equal(ops.host(touch_b), touch)
    for p, (a, b) in enumerate(pairs):
        ma, mb = masks[:, :, a], masks[:, :, b]
        assert inter[p] == np.count_nonzero(ma & mb)
        assert bool(touch[p]) == A.are_mask_connected(ma, mb), (a, b)
    groups = [[0], [1, 2], [3, 4, 5, 6], [9, 1]]
    merged = ops.unpack(ops.union(planes, H, W, groups), H, W).astype(bool)
    for g, members in enumerate(groups):
        want = masks[:, :, members[0]]
        for k in members[1:]:
            want = A.merge_masks(want, masks[:, :, k])
        assert np.array_equal(merged[g], want)


def _shape_masks(rng, H, W, n):
    return C.random_detections(rng, H, W, n)[0]


@pytest.mark.parametrize("H,W,n", [(6, 9, 4), (31, 45, 8), (64, 64, 12), (100, 257, 6), (256, 256, 24)])
def test_labels_match_oracle(ops, H, W, n):
    rng = np.random.default_rng(H + W)
    masks = np.concatenate([_noise_masks(rng, H, W, n // 2), _shape_masks(rng, H, W, n - n // 2)], axis=2)
    # spirals / long snakes: deep union-find chains
    snake = np.zeros((H, W), dtype=bool)
    snake[::2, :] = True
    for r in range(1, H, 2):
        snake[r, (W - 1) if (r // 2) % 2 == 0 else 0] = True
    masks[:, :, 0] = snake
    planes = _planes(ops, masks)
    labels, counts = ops.label(planes, H, W)
    labels, counts = ops.host(labels), ops.host(counts)
    src, comp = [], []
    for i in range(n):
        want, nwant = A.label_components(masks[:, :, i])
        assert counts[i] == nwant, i
        assert np.array_equal(labels[i], want), i
        for c in range(1, min(nwant, 3) + 1):
            src.append(i)
            comp.append(c)
    parts = ops.unpack(ops.select(ops.to_dev(labels, np.int32), H, W, src, comp), H, W)
    for k, (i, c) in enumerate(zip(src, comp)):
        assert np.array_equal(parts[k].astype(bool), labels[i] == c)


def _run_product(masks, class_ids, scores, options, origin=(0, 0), name="img"):
    from mrcnn.analyze import Analyzer

    class Cfg:
        NUM_CLASSES = len(CLASS_NAMES)

    an = Analyzer(None, Cfg())
    an.class_names = CLASS_NAMES
    an.masks, an.boxes, an.class_ids, an.scores = masks, np.zeros((masks.shape[2], 4), np.int32), class_ids, scores
    an.nobjects = masks.shape[2]
    an.image = np.zeros(masks.shape[:2] + (3,), dtype=np.uint8)
    an.image_id = name
    an.image_xmin, an.image_ymin = origin
    an.obj_name_tag = "t0"
    for k, v in options.items():
        setattr(an, k, v)
    an.extract_det_masks()
    an.make_json_results()
    return an


def test_analyzer_matches_reference_goldens():
    golden = C.load_golden()
    assert golden["class_names"] == CLASS_NAMES
    for case in golden["cases"]:
        masks, class_ids, scores = C.case_inputs(case)
        an = _run_product(masks, class_ids, scores, case["options"], tuple(case["origin"]), case["name"])
        from oracle import contours as OC
        for obj, mask in zip(an.results["objs"], an.masks_final):
            # the goldens carry no vertexes (find_contours was stubbed when they were made): check them against the
            # oracle's restatement of the scikit-image algorithm, then blank them for the golden comparison
            assert obj["vertexes"] == OC.mask_vertexes(np.asarray(mask) != 0, xmin=case["origin"][0], ymin=case["origin"][1]), case["name"]
            obj["vertexes"] = []
        got = C.summarise(an.results["objs"], an.masks_final, an.captions, case["H"] * case["W"] <= 64 * 64)
        assert got == case["objs"], (case["name"], case["options"], case["origin"])


@pytest.mark.parametrize("H,W,n,seed", [(48, 80, 30, 1), (128, 128, 60, 2), (256, 256, 100, 3)])
def test_analyzer_matches_oracle_on_random_detections(H, W, n, seed):
    rng = np.random.default_rng(seed)
    masks, class_ids, scores = C.random_detections(rng, H, W, n, density=0.5 if n > 50 else 1.0)
    scores[::7] = scores[3]                    # score ties
    for options in (dict(), dict(split_masks=True), dict(split_masks=True, merge_overlapped_masks=False),
                    dict(split_source_sidelobe=False, merge_overlap_iou_thr=0.05, score_thr=0.6)):
        if n == 100 and options.get("split_masks"):
            continue                           # the oracle needs minutes for the O(N^2) labelling at this size
        an = _run_product(masks, class_ids, scores, options)
        det = A.extract_det_masks(masks, n, class_ids, scores, CLASS_NAMES, **options)
        ref = A.make_json_results(det, CLASS_NAMES, (H, W, 3), image_id="img", obj_name_tag="t0")
        assert len(an.masks_final) == len(det["masks_final"])
        for k in range(len(det["masks_final"])):
            assert np.array_equal(np.asarray(an.masks_final[k]) != 0, np.asarray(det["masks_final"][k]) != 0)
            assert np.asarray(an.masks_final[k]).dtype == np.asarray(det["masks_final"][k]).dtype
            assert np.array_equal(an.bboxes[k], det["bboxes"][k])
        for obj in an.results["objs"]:
            obj["vertexes"] = []
        assert C.summarise(an.results["objs"], an.masks_final, an.captions, True) == \
            C.summarise(ref["objs"], det["masks_final"], det["captions"], True), options


def test_reference_helper_methods(ops):
    from mrcnn.analyze import Analyzer

    class Cfg:
        NUM_CLASSES = 6

    an = Analyzer(None, Cfg())
    rng = np.random.default_rng(9)
    a, b = rng.random((30, 41)) < 0.2, rng.random((30, 41)) < 0.2
    assert np.array_equal(an.merge_masks(a, b), A.merge_masks(a, b)) and an.merge_masks(a, b).dtype == np.bool_
    ai = a.astype(np.int64)
    assert np.array_equal(an.merge_masks(ai, b), A.merge_masks(ai, b)) and an.merge_masks(ai, b).dtype == np.int64
    labels, n = an.extract_mask_connected_components(a)
    want, nwant = A.label_components(a)
    assert n == nwant and np.array_equal(labels, want)
    assert an.are_mask_connected(a, b) == A.are_mask_connected(a, b)
    c = np.zeros_like(a)
    assert an.are_mask_connected(a, c) is False


def test_predict_maps_device_path_matches_host_path():
    """Device-resident batch path (masks never leave the GPU) == per-image host-array path == oracle."""
    import synth
    from mrcnn import model as modellib
    from mrcnn.analyze import Analyzer
    from oracle import network as N
    from test_gpu_engine import _config

    B = 2
    cfg = _config(B)
    cfg.CLASS_NAMES = ["bkg", "spurious", "compact", "extended"]
    m = modellib.MaskRCNN(mode="inference", config=cfg, model_dir="/tmp/mrcnn_logs")
    m.set_weights(N.make_random_weights(0, 4))
    maps = synth.radio_maps(B, 132)
    host = m.detect_maps(np.stack(maps))
    scores = np.concatenate([r["scores"] for r in host])
    assert len(scores) > 4
    thr = float(np.sort(scores)[len(scores) // 3])          # random weights: pick a threshold that keeps most detections
    an = Analyzer(m, cfg)
    an.score_thr = thr
    an.obj_name_tag = "tile"
    batch = an.predict_maps(np.stack(maps), image_ids=["a", "b"])
    assert len(batch) == B and all(isinstance(o["pixels"], np.ndarray) for cat in batch for o in cat["objs"])
    an.pixels_as_lists = True
    batch = an.predict_maps(np.stack(maps), image_ids=["a", "b"])
    total = 0
    for b in range(B):
        r = host[b]
        one = Analyzer(m, cfg)
        one.score_thr, one.obj_name_tag = thr, "tile"
        one.class_names = cfg.CLASS_NAMES
        one.image = np.zeros((132, 132, 3), np.uint8)
        one.image_id = ["a", "b"][b]
        one.masks, one.boxes, one.class_ids, one.scores = r["masks"], r["rois"], r["class_ids"], r["scores"]
        one.extract_det_masks()
        one.make_json_results()
        for res in (one.results, batch[b]):
            for obj in res["objs"]:
                obj["vertexes"] = []
        assert one.results == batch[b]
        det = A.extract_det_masks(np.asarray(r["masks"]), r["rois"].shape[0], r["class_ids"], r["scores"], cfg.CLASS_NAMES,
                                  score_thr=thr)
        ref = A.make_json_results(det, cfg.CLASS_NAMES, (132, 132, 3), image_id=["a", "b"][b], obj_name_tag="tile")
        assert ref == one.results
        total += len(ref["objs"])
    assert total > 0


def test_predict_maps_stream_overlapped_equals_batch_by_batch():
    """Three batches through the overlapped generator (two result slots in flight) == one predict_maps() per batch."""
    import synth
    from mrcnn import model as modellib
    from mrcnn.analyze import Analyzer
    from oracle import network as N
    from test_gpu_engine import _config

    B = 2
    cfg = _config(B)
    cfg.CLASS_NAMES = ["bkg", "spurious", "compact", "extended"]
    m = modellib.MaskRCNN(mode="inference", config=cfg, model_dir="/tmp/mrcnn_logs")
    m.set_weights(N.make_random_weights(0, 4))
    batches = [np.stack(synth.radio_maps(B, 132, start=s)) for s in (0, 40, 80)]
    an = Analyzer(m, cfg)
    an.pixels_as_lists = True
    scores = np.concatenate([r["scores"] for r in m.detect_maps(batches[0])])
    an.score_thr = float(np.sort(scores)[len(scores) // 3])
    want = [an.predict_maps(b, image_ids=["x%d" % k, "y%d" % k], origins=[(0, 0), (7, 3)]) for k, b in enumerate(batches)]
    got = list(an.predict_maps_stream((b, ["x%d" % k, "y%d" % k], [(0, 0), (7, 3)]) for k, b in enumerate(batches)))
    assert got == want
    assert sum(len(c["objs"]) for cats in got for c in cats) > 0
