"""CPU: host-side mirror of the reference interface (mrcnn package) + C-ABI library surface.
No compute call needs a GPU here."""
import json
import os
import re

import numpy as np
import pytest

from oracle import host_ops as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_config_defaults_match_reference(golden):
    from mrcnn.config import Config
    base = Config()
    names = sorted(a for a in dir(base) if not a.startswith("__") and not callable(getattr(base, a)))
    assert names == list(golden["config_attr_names"])
    for n, r in zip(golden["config_attr_names"], golden["config_attr_reprs"]):
        assert repr(getattr(base, n)) == r, n

    class C(Config):
        NUM_CLASSES = 4
        GPU_COUNT = 1
        IMAGES_PER_GPU = 1
        IMAGE_MIN_DIM = 256
        IMAGE_MAX_DIM = 256
    c = C()
    assert [c.BATCH_SIZE, c.IMAGE_META_SIZE] + list(c.IMAGE_SHAPE) == list(golden["config_derived"])
    c.display()


def test_anchor_and_box_helpers_match_reference(golden):
    from mrcnn import utils

    class Cfg:
        BACKBONE = "resnet101"
        BACKBONE_STRIDES = [4, 8, 16, 32, 64]
    for S in (256, 128):
        shapes = utils.compute_backbone_shapes(Cfg, (S, S, 3))
        assert np.array_equal(shapes, golden["backbone_shapes_%d" % S])
        a = utils.generate_pyramid_anchors((4, 8, 16, 32, 64), [0.5, 1, 2], shapes, [4, 8, 16, 32, 64], 1)
        assert np.array_equal(a, golden["anchors_px_%d" % S])
        assert np.array_equal(utils.norm_boxes(a, (S, S)), golden["anchors_norm_%d" % S])
    assert np.array_equal(utils.generate_anchors(32, [0.5, 1, 2], [3, 5], 16, 2), golden["gen_anchors_small"])
    assert np.array_equal(utils.norm_boxes(golden["norm_in"], (132, 200)), golden["norm_out"])
    assert np.array_equal(utils.denorm_boxes(golden["denorm_in"], (132, 132)), golden["denorm_out"])


def test_square_geometry_matches_reference_resize_bookkeeping(golden):
    from mrcnn import utils
    h, w = golden["resize_in"].shape[:2]
    scale, out_hw, top_left, window, padding = utils.square_geometry(h, w, 128, 128, 0, "square")
    assert float(scale) == float(golden["resize_scale"][0])
    assert tuple(window) == tuple(golden["resize_window"])
    assert np.array_equal(np.array(padding), golden["resize_padding"])
    # galaxy0002 case of SURVEY.md Appendix D: 132 -> 256, no padding
    scale, out_hw, top_left, window, _ = utils.square_geometry(132, 132, 256, 256, 0, "square")
    assert out_hw == (256, 256) and window == (0, 0, 256, 256) and abs(scale - 256 / 132) < 1e-12
    # oracle agreement on a down-scaling case
    img = np.zeros((300, 500, 3), np.uint8)
    _, owin, oscale, opad, _ = H.resize_image(img, min_dim=256, max_dim=256, min_scale=0, mode="square")
    scale, out_hw, top_left, window, padding = utils.square_geometry(300, 500, 256, 256, 0, "square")
    assert (scale, tuple(window), padding) == (oscale, tuple(owin), opad)


def test_meta_helpers(golden):
    from mrcnn import model as modellib
    meta = modellib.compose_image_meta(3, (132, 132, 3), (256, 256, 3), (0, 0, 256, 256), 1.9393939,
                                       np.zeros([4], dtype=np.int32))
    assert np.array_equal(meta, golden["compose_meta"])
    p = modellib.parse_image_meta(meta[None])
    assert p["image_shape"].tolist() == [[256, 256, 3]] and p["window"].tolist() == [[0, 0, 256, 256]]

    class Cfg:
        MEAN_PIXEL = np.array([0, 0, 0])
    assert np.array_equal(modellib.mold_image(golden["resize_in"][:4, :4], Cfg), golden["mold_image_f"])


def test_fits_reader(golden_dir, tmp_path):
    from mrcnn import fitsio
    for name in ("galaxy0002.fits", "sidelobe0001.fits"):
        path = os.path.join(golden_dir, name)
        data, hdr = fitsio.read_primary(path)
        ref, rhdr = H.parse_fits_primary(open(path, "rb").read())
        assert data.dtype == np.float32 and data.dtype.isnative
        assert np.array_equal(data, ref, equal_nan=True)
        assert hdr["NAXIS1"] == 132 and hdr["BITPIX"] == -32
    _, hdr = fitsio.read_primary(os.path.join(golden_dir, "galaxy0002.fits"))
    assert hdr["TELESCOP"] == "EVLA" and abs(hdr["BMAJ"] - 1.7778e-3) < 1e-6
    # 4-D cube + int16 with BSCALE/BZERO, written by hand
    cards = ["SIMPLE  =                    T", "BITPIX  =                   16", "NAXIS   =                    4",
             "NAXIS1  =                    3", "NAXIS2  =                    2", "NAXIS3  =                    1",
             "NAXIS4  =                    1", "BSCALE  =                  0.5", "BZERO   =                 10.0", "END"]
    hdrb = "".join(c.ljust(80) for c in cards).ljust(2880).encode()
    payload = np.arange(6, dtype=">i2").tobytes().ljust(2880, b"\0")
    f = tmp_path / "cube.fits"
    f.write_bytes(hdrb + payload)
    data, hdr = fitsio.read_primary(str(f))
    assert data.shape == (1, 1, 2, 3) and np.allclose(data.ravel(), np.arange(6) * 0.5 + 10)
    with pytest.raises(IOError):
        fitsio.read_primary(__file__)


def test_c_abi_library_exports_every_declared_symbol():
    import ctypes
    from mrcnn import _native
    hdr = open(os.path.join(ROOT, "include", "mrcnn_b200.h")).read()
    declared = set(re.findall(r"\b(mrcnn_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_native.SIGNATURES), declared ^ set(_native.SIGNATURES)
    lib = _native.lib()                       # loads without a GPU (static cudart, no driver needed)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.mrcnn_abi_version() == 2
    assert ctypes.sizeof(_native.ConvDesc) == 14 * 4
    assert ctypes.sizeof(_native.EngineConfig) == (11 + 3 + 8 + 5) * 4


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import ctypes
    from mrcnn import _native, model as modellib
    from mrcnn.config import Config

    class C(Config):
        NUM_CLASSES = 4
        IMAGE_MIN_DIM = 256
        IMAGE_MAX_DIM = 256
    with pytest.raises(_native.NativeError, match="no CPU fallback"):
        modellib.MaskRCNN(mode="inference", config=C(), model_dir="/tmp/x")
    lib = _native.lib()
    cfg = _native.EngineConfig(batch_size=1, image_size=256, num_classes=4, pre_nms_limit=6000, post_nms_rois=1000,
                               detection_max_instances=100, pool_size=7, mask_pool_size=14, fc_layers_size=1024,
                               top_down_pyramid_size=256, anchors_per_location=3)
    for i, s in enumerate((4, 8, 16, 32, 64)):
        cfg.backbone_strides[i] = s
    h = ctypes.c_void_p()
    assert lib.mrcnn_engine_create(ctypes.byref(cfg), 0, ctypes.byref(h)) != 0
    assert b"CUDA" in lib.mrcnn_last_error() or b"device" in lib.mrcnn_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "caesar-mrcnn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), os.path.join(dirpath, f)
                assert "oracle/" not in src or f.endswith((".cu", ".cuh")), os.path.join(dirpath, f)


def test_result_buffer_sets_return_to_the_pool_when_the_last_view_dies():
    """detect() hands out numpy views of pinned buffer sets; a set is recycled only after every view is gone."""
    import gc
    import torch
    from mrcnn import model as M
    pool = {}
    pset = M._PinnedSet(torch, 2, 3, 4, 5, dw=1, pin=False)
    bufs = M._lease_arrays(pool, pset)
    assert [b.shape for b in bufs[:6]] == [(2, 3, 4), (2, 3), (2, 3), (2,), (2, 20, 1), (2, 60)] and bufs[6] == (4, 5)
    assert [b.dtype for b in bufs[:6]] == [np.int32, np.int32, np.float32, np.int32, np.uint32, np.uint8]
    bufs[4][:] = 5                                                  # every pixel: detections 0 and 2
    assert int(pset.tensors[4].sum()) == 5 * 2 * 20                 # same memory
    bufs[3][:] = [1, 2]
    res = M.MaskRCNN._results_from_buffers(bufs)                    # host-only C++ expansion, no GPU needed
    assert res[1]["masks"].shape == (4, 5, 2) and res[1]["masks"].dtype == np.bool_
    assert res[1]["masks"].flags["C_CONTIGUOUS"] and res[0]["masks"].shape == (4, 5, 1)
    assert res[1]["masks"][:, :, 0].all() and not res[1]["masks"][:, :, 1].any() and res[0]["masks"].all()
    keep = res[1]["masks"]
    del bufs, res
    gc.collect()
    assert not pool.get(pset.key)                                   # one view still alive
    del keep
    gc.collect()
    assert pool[pset.key] == [pset]


def test_host_expand_mask_bits_matches_numpy():
    """mrcnn_host_expand_mask_bits (host-only C++, SIMD and portable paths, 1..n threads) against a numpy unpack:
    ragged detection counts incl. 0, 1, 63..65 and the maximum, odd pixel counts."""
    import ctypes
    import subprocess
    import sys
    from mrcnn import _native
    lib = _native.lib()
    rng = np.random.default_rng(3)
    for D, npx in ((100, 4099), (64, 257), (256, 1000), (7, 33)):
        dw = lib.mrcnn_mask_bits_words(D)
        assert dw * 32 >= D and (dw & (dw - 1)) == 0
        counts = np.array([0, 1, min(D, 63), min(D, 64), min(D, 65), D, int(rng.integers(1, D + 1))], dtype=np.int32)
        B = len(counts)
        bits = rng.integers(0, 2 ** 32, size=(B, npx, dw), dtype=np.uint64).astype(np.uint32)
        want = np.unpackbits(bits.view(np.uint8).reshape(B, npx, dw * 4), axis=2, bitorder="little")
        for threads in (1, 3, 0):
            dense = np.full((B, npx * D + 64), 0xAB, dtype=np.uint8)
            dst = (ctypes.c_void_p * B)(*[dense[i].ctypes.data for i in range(B)])
            _native.check(lib.mrcnn_host_expand_mask_bits(bits.ctypes.data, B, npx, dw, counts.ctypes.data, dst, threads))
            for i, n in enumerate(counts):
                assert np.array_equal(dense[i, :npx * n].reshape(npx, n), want[i, :, :n]), (D, npx, i, n, threads)
                assert (dense[i, npx * n:] == 0xAB).all(), "wrote past the image's [npx, n] block"
    # the portable path (what a CPU without AVX-512BW runs) in a fresh process
    code = ("import os, sys, ctypes, numpy as np; sys.path.insert(0, %r); from mrcnn import _native; lib = _native.lib();"
            "bits = (np.arange(3 * 50 * 4, dtype=np.uint64).reshape(3, 50, 4) * 2654435761 %% 2**32).astype(np.uint32);"
            "counts = np.array([100, 37, 8], dtype=np.int32); dense = np.zeros((3, 5000), np.uint8);"
            "dst = (ctypes.c_void_p * 3)(*[dense[i].ctypes.data for i in range(3)]);"
            "assert lib.mrcnn_host_expand_mask_bits(bits.ctypes.data, 3, 50, 4, counts.ctypes.data, dst, 2) == 0;"
            "want = np.unpackbits(bits.view(np.uint8).reshape(3, 50, 16), axis=2, bitorder='little');"
            "assert all(np.array_equal(dense[i, :50 * n].reshape(50, n), want[i, :, :n]) for i, n in enumerate(counts)); print('ok')"
            % os.path.join(ROOT, "caesar-mrcnn_b200"))
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, MRCNN_B200_HOST_SIMD="0"), capture_output=True, text=True)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr


def test_analyzer_host_logic_matches_oracle_graph_and_iou():
    """Host pieces of mrcnn/analyze.py that need no GPU: DFS component order, pair enumeration, IOU arithmetic."""
    from mrcnn import analyze as P
    from oracle import analyze_ops as A
    rng = np.random.default_rng(11)
    for _ in range(50):
        n = int(rng.integers(1, 14))
        edges = [(int(a), int(b)) for a, b in rng.integers(0, n, size=(int(rng.integers(0, 2 * n)), 2)) if a != b]
        gp, go = P.Graph(n), A.Graph(n)
        for v, w in edges:
            gp.addEdge(v, w)
            go.add_edge(v, w)
        assert gp.connectedComponents() == go.connected_components()
    pairs, slices = P._all_pairs([3, 0, 1, 4])
    assert pairs.tolist() == [[0, 1], [0, 2], [1, 2], [4, 5], [4, 6], [4, 7], [5, 6], [5, 7], [6, 7]]
    assert slices == [(0, 3, 0), (3, 3, 3), (3, 3, 3), (3, 9, 4)]
    a = rng.random((40, 33)) < 0.3
    b = rng.random((40, 33)) < 0.3
    z = np.zeros_like(a)
    inter = np.array([np.count_nonzero(a & b), 0, 0], dtype=np.int32)
    got = P._iou(inter, np.array([a.sum(), a.sum(), 0], np.int32), np.array([b.sum(), 0, 0], np.int32))
    assert got.dtype == np.float64
    assert got.tolist() == [float(A.jaccard_binary(a, b)), float(A.jaccard_binary(a, z)), float(A.jaccard_binary(z, z))]


def test_analyzer_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mrcnn.analyze import Analyzer

    class Cfg:
        NUM_CLASSES = 4

    an = Analyzer(None, Cfg())
    an.class_names = ["bkg", "a", "b", "c"]
    an.masks = np.ones((8, 8, 1), dtype=bool)
    an.boxes, an.class_ids, an.scores = np.zeros((1, 4), np.int32), np.array([1], np.int32), np.array([0.9], np.float32)
    with pytest.raises(Exception):
        an.extract_det_masks()


def test_analyzer_host_pipeline_matches_oracle_with_numpy_backend():
    """analyze_frames (the product's host logic: ordering, merge graph, cliques, selection, result assembly) driven
    by a numpy test double of the device primitives == the oracle, for several frames at once."""
    import analyzer_cases as C
    from mrcnn import analyze as P
    from oracle import analyze_ops as A
    names = ["bkg", "spurious", "compact", "extended", "extended-multisland", "flagged"]
    rng = np.random.default_rng(1)
    F, S, D = 4, 48, 24
    masks = np.zeros((F, S, S, D), bool)
    cls, sc = np.zeros((F, D), np.int32), np.zeros((F, D), np.float32)
    for f in range(F):
        masks[f], cls[f], sc[f] = C.random_detections(rng, S, S, D, density=0.6)
    masks[2] = False                                     # a frame whose detections are all empty masks
    sc[1, ::3] = sc[1, 0]                                # score ties
    ops = C.NumpyPlaneOps(masks)
    frames = [P._Frame(4096 + f * S * S * D, D, D, cls[f], sc[f]) for f in range(F)]
    origins = [(0, 0), (5, 9), (0, 0), (100, 200)]
    for opts in (dict(), dict(split_source_sidelobe=False, merge_overlap_iou_thr=0.05, score_thr=0.6),
                 dict(merge_overlapped_masks=False), dict(select_best_overlapped_masks=False)):
        res = P.analyze_frames(ops, frames, S, S, names, origins=origins, want_masks=True, **opts)
        for f in range(F):
            det = A.extract_det_masks(masks[f], D, cls[f], sc[f], names, **opts)
            ref = A.make_json_results(det, names, (S, S, 3), xmin=origins[f][1], ymin=origins[f][0])
            got = res[f]
            assert len(ref["objs"]) == len(got.class_ids_final), (f, opts)
            for k, o in enumerate(ref["objs"]):
                assert o["pixels"] == got.pixels[k].tolist()
                assert o["score"] == got.scores_final[k] and type(o["score"]) is type(got.scores_final[k])
                assert o["class_id"] == int(got.class_ids_final[k])
                assert list(det["bboxes"][k]) == list(got.bboxes[k])
                assert det["captions"][k] == got.captions[k]
                assert np.array_equal(np.asarray(det["masks_final"][k]) != 0, got.masks_final[k] != 0)


def test_analyzer_rejects_drawing_before_any_work_and_handles_no_detections():
    from mrcnn.analyze import Analyzer

    class Cfg:
        NUM_CLASSES = 4

    class Boom:
        def detect(self, images, verbose=0):
            raise AssertionError("detect must not be reached")

    an = Analyzer(Boom(), Cfg())
    an.draw = True
    with pytest.raises(NotImplementedError):
        an.predict(np.zeros((8, 8, 3), np.uint8), "img")
    an = Analyzer(None, Cfg())                       # no detections: no device work, empty lists
    an.class_names = ["bkg", "a", "b", "c"]
    an.masks, an.boxes = np.empty((8, 8, 0)), np.zeros((0, 4), np.int32)
    an.class_ids, an.scores = np.zeros((0,), np.int32), np.zeros((0,), np.float32)
    an.image = np.zeros((8, 8, 3), np.uint8)
    an.extract_det_masks()
    an.make_json_results()
    assert an.masks_final == [] and an.bboxes == [] and an.results == {"image_id": -1, "objs": []}


def test_fits_header_only_memmap_subimage_and_writer(golden_dir, tmp_path):
    """Survey-sized access paths of mrcnn/fitsio.py against read_primary on the shipped files and on written ones."""
    from mrcnn import fitsio, utils
    for name in ("galaxy0002.fits", "sidelobe0001.fits"):
        path = os.path.join(golden_dir, name)
        full, hdr = fitsio.read_primary(path)
        hdr2, pos = fitsio.read_header_only(path)
        assert dict(hdr2) == dict(hdr) and pos % 2880 == 0
        plane = full[0, 0] if full.ndim == 4 else full
        sub, _ = fitsio.read_subimage(path, 5, 77, 11, 40)
        assert sub.dtype == plane.dtype and np.array_equal(sub, plane[11:40, 5:77], equal_nan=True)
        mm, _ = fitsio.open_primary(path)
        assert mm.shape == full.shape and not mm.flags.writeable
        assert utils.get_fits_size(path) == (hdr["NAXIS1"], hdr["NAXIS2"])
    rng = np.random.default_rng(3)
    cases = [(rng.normal(size=(37, 53)).astype(np.float32), ()), (rng.normal(size=(1, 1, 20, 31)), ()),
             (rng.integers(-3000, 3000, size=(25, 40)).astype(np.int16), (("BSCALE", 0.5), ("BZERO", 100.0), ("BLANK", -7))),
             (rng.integers(0, 255, size=(16, 16)).astype(np.uint8), ()), (rng.integers(-9, 9, size=(9, 70)).astype(np.int32), ())]
    for k, (arr, cards) in enumerate(cases):
        path = str(tmp_path / ("w%d.fits" % k))
        if cards:
            arr[3, 4] = -7
        fitsio.write_primary(path, arr, cards + (("OBJECT", "it's a test"), ("BMAJ", 1.25e-3)))
        full, hdr = fitsio.read_primary(path)
        assert hdr["OBJECT"] == "it's a test" and hdr["BMAJ"] == 1.25e-3 and hdr["NAXIS"] == arr.ndim
        if cards:
            want = arr.astype(np.float64) * 0.5 + 100.0
            want[arr == -7] = np.nan
            assert np.array_equal(full, want.astype(np.float32), equal_nan=True)
        else:
            assert np.array_equal(full, arr) and full.dtype == arr.dtype
        plane = full[0, 0] if full.ndim == 4 else full
        sub, _ = fitsio.read_subimage(path, 2, 13, 1, 8)
        assert np.array_equal(sub, plane[1:8, 2:13], equal_nan=True) and sub.dtype == plane.dtype
    with pytest.raises(fitsio.FitsError):
        fitsio.read_header_only(__file__)
    with pytest.raises(fitsio.FitsError):
        fitsio.write_primary(str(tmp_path / "bad.fits"), np.zeros((3,), np.float32))


def test_analyzer_host_logic_reproduces_reference_goldens_with_numpy_backend():
    """The reference Analyzer's own outputs (tests/golden/analyzer_golden.json) replayed through the product's host
    logic with the numpy test double in place of the device primitives (scipy labelling stands in for the device's
    4-connected labelling in the split_masks cases)."""
    import analyzer_cases as C
    from mrcnn import analyze as P
    golden = C.load_golden()
    names = golden["class_names"]
    done = 0
    for case in golden["cases"]:
        masks, class_ids, scores = C.case_inputs(case)
        H, W, D = masks.shape
        ops = C.NumpyPlaneOps(masks[None])
        frame = P._Frame(4096, D, D, class_ids, scores)
        xmin, ymin = case["origin"]
        res = P.analyze_frames(ops, [frame], H, W, names, origins=[(ymin, xmin)], want_masks=True, **case["options"])[0]
        cat = P.build_json_results(case["name"], "t0", names, H, W, xmin, ymin, res.masks_final, res.class_ids_final,
                                   res.scores_final, res.bboxes, res.pixels)
        for obj in cat["objs"]:
            obj["vertexes"] = []
        got = C.summarise(cat["objs"], res.masks_final, res.captions, H * W <= 64 * 64)
        want = case["objs"]                      # mask dtypes included: int64 for masks that went through the component split
        assert got == want, (case["name"], case["options"], case["origin"])
        done += 1
    assert done >= 50


def test_native_merge_components_reproduces_the_reference_graph_order():
    """mrcnn_host_merge_components (host-only C++) == mrcnn/graph.py semantics (oracle Graph) on random multi-frame
    graphs: component order, member pre-order, components per frame."""
    import ctypes
    from mrcnn import _native, analyze as P
    from oracle import analyze_ops as A
    lib = _native.lib()
    rng = np.random.default_rng(21)
    for trial in range(60):
        F = int(rng.integers(1, 6))
        counts = [int(rng.integers(0, 15)) for _ in range(F)]
        pairs, _ = P._all_pairs(counts)
        mergeable = (rng.random(len(pairs)) < rng.choice([0.0, 0.05, 0.3, 1.0])).astype(np.uint8)
        n = sum(counts)
        members = np.full(n, -1, np.int32)
        offsets = np.full(n + 1, -1, np.int32)
        frame_comps = np.zeros(F, np.int32)
        ncomp = ctypes.c_int32(-1)
        counts_arr = np.asarray(counts, np.int32)
        _native.check(lib.mrcnn_host_merge_components(F, counts_arr.ctypes.data, pairs.ctypes.data if len(pairs) else None,
                                                      mergeable.ctypes.data if len(pairs) else None, len(pairs),
                                                      members.ctypes.data, offsets.ctypes.data, frame_comps.ctypes.data,
                                                      ctypes.byref(ncomp)), "merge_components")
        want, want_frames, base, pos = [], [], 0, 0
        for c in counts:
            g = A.Graph(c)
            npairs = c * (c - 1) // 2
            for k in np.nonzero(mergeable[pos:pos + npairs])[0]:
                g.add_edge(int(pairs[pos + k, 0]) - base, int(pairs[pos + k, 1]) - base)
            cc = g.connected_components()
            want += [[base + v for v in comp] for comp in cc]
            want_frames.append(len(cc))
            base += c
            pos += npairs
        got = [members[offsets[i]:offsets[i + 1]].tolist() for i in range(ncomp.value)]
        assert got == want and frame_comps.tolist() == want_frames and ncomp.value == len(want)


def test_host_pair_flags_and_pair_lists_match_the_numpy_expressions():
    """mrcnn_host_all_pairs / mrcnn_host_pair_flags (host C++) against the numpy expressions of the per-frame walk:
    pair order, merge edges (connected, same class, float64 IOU >= thr), selection edges (spurious-vs-source rule), the
    'a linked mask scores higher' flags and the per-frame tie detector — including empty unions and equal scores."""
    from mrcnn import analyze as P
    rng = np.random.default_rng(3)
    for trial in range(40):
        F = int(rng.integers(1, 7))
        counts = rng.integers(0, 12, size=F).astype(np.int32)
        n = int(counts.sum())
        c2, pairs = P._host_pairs(counts)
        want_pairs, _ = P._all_pairs(counts.tolist())
        assert np.array_equal(pairs, want_pairs.reshape(-1, 2)) and np.array_equal(c2, counts)
        npairs = len(pairs)
        area = rng.integers(0, 50, size=n).astype(np.int32)
        inter = np.minimum(rng.integers(0, 50, size=npairs), np.minimum(area[pairs[:, 0]], area[pairs[:, 1]])).astype(np.int32) if npairs else np.zeros(0, np.int32)
        touch = (rng.random(npairs) < 0.6).astype(np.int32)
        cls = rng.integers(1, 4, size=n).astype(np.int32)
        score = rng.choice(np.array([0.5, 0.625, 0.75, 0.875], np.float32), size=n)
        thr = float(rng.choice([0.0, 0.3, 0.5]))
        iou = P._iou(inter, area[pairs[:, 0]], area[pairs[:, 1]]) if npairs else np.zeros(0)
        flags, _, _ = P._pair_flags(0, counts, cls, None, area, inter, touch, True, thr, n)
        want = (touch != 0) & (cls[pairs[:, 0]] == cls[pairs[:, 1]]) & (iou >= thr) if npairs else np.zeros(0, bool)
        assert np.array_equal(flags.astype(bool), want)
        spur = (cls == 1).astype(np.int32)
        for use_iou in (True, False):
            linked, loses, tie = P._pair_flags(1, counts, spur, score, area, inter, touch, use_iou, thr, n)
            want = touch != 0
            if use_iou and npairs:
                want &= ~((spur[pairs[:, 0]] != spur[pairs[:, 1]]) & (iou < thr))
            assert np.array_equal(linked.astype(bool), want)
            li, lj = pairs[want, 0], pairs[want, 1]
            w_loses = np.zeros(n, bool)
            w_loses[li[score[li] < score[lj]]] = True
            w_loses[lj[score[lj] < score[li]]] = True
            assert np.array_equal(loses.astype(bool), w_loses)
            frame_of = np.repeat(np.arange(F), counts)
            w_tie = np.zeros(F, bool)
            w_tie[frame_of[li[score[li] == score[lj]]]] = True
            assert np.array_equal(tie.astype(bool), w_tie)


def test_generic_walk_and_array_path_give_the_same_catalogues(monkeypatch):
    """The per-frame Python walk (MRCNN_B200_ANALYZE_GENERIC=1, also the split_masks route) == the default array path
    (pair tests, merge components and pair lists in host C++): the reference goldens and the random multi-frame
    comparison against the oracle are replayed with the array path switched off, and a batch with score ties, merges,
    empty frames and per-frame origins goes through both."""
    import analyzer_cases as C
    from mrcnn import analyze as P
    calls = []
    real = P._analyze_batch_arrays
    monkeypatch.setattr(P, "_analyze_batch_arrays", lambda *a, **k: (calls.append(1), real(*a, **k))[1])
    test_analyzer_host_logic_reproduces_reference_goldens_with_numpy_backend()
    assert len(calls) >= 30                              # the default route IS the array path
    monkeypatch.setattr(P, "_FORCE_GENERIC", True)
    n = len(calls)
    test_analyzer_host_logic_reproduces_reference_goldens_with_numpy_backend()
    test_analyzer_host_pipeline_matches_oracle_with_numpy_backend()
    assert len(calls) == n

    names = ["bkg", "spurious", "compact", "extended", "extended-multisland", "flagged"]
    rng = np.random.default_rng(5)
    F, S, D = 6, 40, 30
    masks = np.zeros((F, S, S, D), bool)
    cls, sc = np.zeros((F, D), np.int32), np.zeros((F, D), np.float32)
    for f in range(F):
        masks[f], cls[f], sc[f] = C.random_detections(rng, S, S, D, density=0.7)
    masks[3] = False
    sc[1, ::2] = sc[1, 0]                                # ties between linked masks -> clique route
    sc[4, :] = 0.1                                       # a frame with nothing above the threshold
    cls[5, :] = 2                                        # one class: long merge chains
    ops = C.NumpyPlaneOps(masks)
    frames = [P._Frame(4096 + f * S * S * D, D, D - (f == 2) * 7, cls[f], sc[f]) for f in range(F)]
    origins = [(0, 0), (5, 9), (0, 0), (100, 200), (3, 3), (7, 0)]
    for opts in (dict(), dict(split_source_sidelobe=False, merge_overlap_iou_thr=0.05, score_thr=0.6),
                 dict(merge_overlapped_masks=False), dict(select_best_overlapped_masks=False), dict(score_thr=2.0),
                 dict(split_masks=True), dict(split_masks=True, merge_overlapped_masks=False),
                 dict(split_masks=True, merge_overlap_iou_thr=0.02, split_source_sidelobe=False)):
        res = {}
        for generic in (False, True):
            monkeypatch.setattr(P, "_FORCE_GENERIC", generic)
            res[generic] = P.analyze_frames(ops, frames, S, S, names, origins=origins, want_masks=True, vertexes="lists", **opts)
        for f in range(F):
            a, g = res[False][f], res[True][f]
            assert len(a.class_ids_final) == len(g.class_ids_final)
            assert [int(c) for c in a.class_ids_final] == [int(c) for c in g.class_ids_final]
            assert [(type(x), float(x)) for x in a.scores_final] == [(type(x), float(x)) for x in g.scores_final]
            assert a.class_names_final == g.class_names_final and a.captions == g.captions
            assert [b.tolist() for b in a.bboxes] == [b.tolist() for b in g.bboxes]
            assert [p_.tolist() for p_ in a.pixels] == [p_.tolist() for p_ in g.pixels]
            assert all(np.array_equal(x != 0, y != 0) and x.dtype == y.dtype for x, y in zip(a.masks_final, g.masks_final))
            cat_a = P.build_json_results(f, "t", names, S, S, origins[f][1], origins[f][0], a.masks_final, a.class_ids_final,
                                         a.scores_final, a.bboxes, a.pixels, vertexes=a.vertexes)
            cat_g = P.build_json_results(f, "t", names, S, S, origins[f][1], origins[f][0], g.masks_final, g.class_ids_final,
                                         g.scores_final, g.bboxes, g.pixels, vertexes=g.vertexes)
            assert json.dumps(cat_a, cls=P.NumpyEncoder, sort_keys=True) == json.dumps(cat_g, cls=P.NumpyEncoder, sort_keys=True)
    # the batched catalogue builder == build_json_results frame by frame
    monkeypatch.setattr(P, "_FORCE_GENERIC", False)
    batch = real(ops, frames, S, S, names, origins, False, 0.7, True, True, True, 0.3, None, "lists")
    cats = P.build_json_results_batch(batch, ["im%d" % f for f in range(F)], ["t%d" % f for f in range(F)], names, S, S, origins, True)
    per_frame = batch.frame_results(names)
    for f in range(F):
        r = per_frame[f]
        want = P.build_json_results("im%d" % f, "t%d" % f, names, S, S, origins[f][1], origins[f][0], r.masks_final,
                                    r.class_ids_final, r.scores_final, r.bboxes, r.pixels, vertexes=r.vertexes)
        assert json.dumps(cats[f], cls=P.NumpyEncoder, sort_keys=True) == json.dumps(want, cls=P.NumpyEncoder, sort_keys=True)
    assert sum(len(c["objs"]) for c in cats) > 20


def test_drop_in_exports_of_survey_8b_exist():
    """SURVEY.md §8(b) 'must export' list: code written against the reference imports these names."""
    from mrcnn import model as M, utils as U
    for name in ("read_fits", "get_fits_header", "resize_image", "resize", "norm_boxes", "denorm_boxes",
                 "generate_pyramid_anchors", "unmold_mask", "extract_bboxes", "generate_tiles", "Dataset"):
        assert hasattr(U, name), name
    for name in ("MaskRCNN", "mold_image", "unmold_image", "compose_image_meta", "parse_image_meta", "load_image_gt",
                 "data_generator", "build_rpn_targets"):
        assert hasattr(M, name), name
    for name in ("compute_overlaps", "box_refinement", "resize_mask", "minimize_mask", "expand_mask", "trim_zeros",
                 "compute_iou", "compute_overlaps_masks", "non_max_suppression", "apply_box_deltas", "compute_matches",
                 "compute_ap", "compute_ap_range", "compute_recall", "get_iou"):       # pinned: tests/test_utils_extra_golden.py
        assert hasattr(U, name), name
    # module layout of the reference package: the imports at the top of its scripts resolve (scripts/run.py:42-49)
    from mrcnn import logger, visualize                                  # noqa: F401
    from mrcnn.analyze import Analyzer, ModelTester                      # noqa: F401
    from mrcnn.graph import Graph
    from mrcnn.parallel_model import ParallelModel
    from mrcnn.sfinder import SFinder                                    # noqa: F401
    g = Graph(4)
    g.addEdge(0, 2)
    g.addEdge(2, 3)
    assert g.connectedComponents() == [[0, 2, 3], [1]]
    for broken in (lambda: ParallelModel(None, 2), lambda: ModelTester(None, None, None), lambda: visualize.display_instances()):
        with pytest.raises(NotImplementedError):
            broken()
    # the training-path functions are real since round 2 (tests/test_training_host.py pins them to the reference)
    with pytest.raises(NotImplementedError):
        M.load_image_gt(None, None, 0, augmentation=object())


def test_extract_bboxes_matches_reference_loop():
    """utils.extract_bboxes (vectorised) against the reference's per-instance loop (mrcnn/utils.py:49-77) restated in
    oracle/analyze_ops.py: random masks, empty instances, single pixels, full frames, zero instances."""
    from mrcnn import utils as U
    from oracle import analyze_ops as A
    rng = np.random.default_rng(2)
    for H_, W_, n in ((17, 23, 6), (1, 9, 3), (32, 32, 0), (8, 5, 4)):
        m = np.zeros((H_, W_, n), dtype=np.uint8)
        for i in range(n):
            if i % 3 == 0:
                continue                                     # empty instance -> zeros
            y1, x1 = int(rng.integers(0, H_)), int(rng.integers(0, W_))
            y2, x2 = int(rng.integers(y1, H_)) + 1, int(rng.integers(x1, W_)) + 1
            m[y1:y2, x1:x2, i] = rng.integers(0, 2, (y2 - y1, x2 - x1))
            m[y1, x1, i] = 1
        got = U.extract_bboxes(m)
        assert got.dtype == np.int32 and got.shape == (n, 4)
        want = np.array([A.extract_bbox(m[:, :, i]) for i in range(n)], dtype=np.int32).reshape(n, 4)
        assert np.array_equal(got, want), (H_, W_, n)
    full = np.ones((4, 6, 1), bool)
    assert U.extract_bboxes(full).tolist() == [[0, 0, 4, 6]]


def test_dataset_bookkeeping():
    from mrcnn import utils as U
    ds = U.Dataset()
    ds.add_class("rg", 1, "sidelobe")
    ds.add_class("rg", 2, "source, compact")
    ds.add_class("rg", 1, "duplicate ignored")
    ds.add_image("rg", "img7", "/data/a.fits", extra=3)
    ds.prepare()
    assert ds.num_classes == 3 and ds.class_names == ["BG", "sidelobe", "source"] and ds.num_images == 1
    assert ds.map_source_class_id("rg.2") == 2 and ds.get_source_class_id(2, "rg") == 2
    assert ds.image_from_source_map == {"rg.img7": 0} and ds.source_image_link(0) == "/data/a.fits"
    assert sorted(ds.source_class_ids["rg"]) == [0, 1, 2] and ds.image_info[0]["extra"] == 3
    mask, ids = ds.load_mask(0)
    assert mask.shape == (0, 0, 0) and ids.shape == (0,)


def test_graph_limits_are_reported_at_build_time():
    from mrcnn import model as M
    from mrcnn.config import Config

    class Big(Config):
        NUM_CLASSES = 9
        DETECTION_MAX_INSTANCES = 300
    msgs = M.graph_limit_problems(Big())
    assert len(msgs) == 2 and "NUM_CLASSES=9" in msgs[0] and "DETECTION_MAX_INSTANCES=300" in msgs[1]
    assert M.graph_limit_problems(Config()) == []


def test_dense_share_policy_is_opt_in_and_follows_the_callers_waits(monkeypatch):
    """Hybrid delivery of the dense masks (mrcnn/model.py: _dense_share / _balance_dense_share): off unless
    MRCNN_B200_DENSE_SHARE is set; a number fixes the share; `auto` grows it while result() finds its batch already on the
    host (host-bound) and shrinks it while the caller waits for the device or for the DMA copy; never above B - 4."""
    from mrcnn import model as M

    class Dummy:
        _dense_share = M.MaskRCNN._dense_share
        _balance_dense_share = M.MaskRCNN._balance_dense_share

        def __init__(self):
            self._dense_share_state = None

    B, D = 64, 100
    monkeypatch.delenv("MRCNN_B200_DENSE_SHARE", raising=False)
    d = Dummy()
    assert d._dense_share(B, D) == 0 and d._dense_share_state["fixed"] and not M._dense_share_requested()
    monkeypatch.setenv("MRCNN_B200_DENSE_SHARE", "24")
    d = Dummy()
    assert d._dense_share(B, D) == 24 and M._dense_share_requested()
    d._balance_dense_share(B, 24, 5.0, 0, 0.0, 0.0)
    assert d._dense_share(B, D) == 24                              # fixed: never adapted
    assert Dummy()._dense_share(B, 102) == 0                       # slots must be 4-byte multiples
    monkeypatch.setenv("MRCNN_B200_DENSE_SHARE", "auto")
    d = Dummy()
    assert d._dense_share(B, D) == 16 and not d._dense_share_state["fixed"]
    for _ in range(40):                                            # host-bound: everything had arrived when asked
        d._balance_dense_share(B, d._dense_share(B, D), 20.0, 0, 0.01, 0.0)
    assert d._dense_share(B, D) == B - 4
    for _ in range(10):                                            # waiting for the device: hand work back to the cores
        d._balance_dense_share(B, d._dense_share(B, D), 2.0, 0, 3.0, 0.0)
    assert d._dense_share(B, D) == B - 14
    for _ in range(100):                                           # the DMA copy lags behind the cores
        d._balance_dense_share(B, d._dense_share(B, D), 2.0, 0, 0.0, 1.0)
    assert d._dense_share(B, D) == 0
    d._balance_dense_share(B, 0, 2.0, 0, None, 0.0)                # blocking call: no waiting pattern to learn from
    assert d._dense_share(B, D) == 0
