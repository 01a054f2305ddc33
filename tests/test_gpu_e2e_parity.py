"""End-to-end parity of the bf16 sm_100a engine against the fp32 CPU oracle on BASELINE.json configs[0] (galaxy0002.fits)
and configs[1] (synthetic radio maps): FINAL detections (reference outputs mrcnn/model.py:2623-2704 `detect`, taps :2156-2158),
not per-stage tensors.  Detections are matched by class + box IoU (parity_metrics.py); asserted are the matched fraction and
the distribution of |d box| (pixels of the molded frame), |d score| and full-frame mask IoU over the matched pairs.

Tolerances (measured, DESIGN.md §2 "End-to-end floating-point parity"): the north star's example figures (1e-3 px, 1e-3,
IoU >= 0.99) are NOT reachable with bf16 operands through 104 backbone layers and are not claimed; with seeded random
weights (every ROI scored ~0.3-0.5, NMS decisions on a knife edge) the engine's arithmetic, restated on the CPU
(tools/e2e_parity_cpu.py, oracle emulate_bf16=True vs False), gives matched 0.95, |d box| median 0.011 / p95 0.023 px,
|d score| median 2.5e-3 / p95 6.8e-3, mask IoU median 1.0 with 96 % of the pairs >= 0.9.  The bounds below are ~2x that.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

import parity_metrics as PM  # noqa: E402
import synth  # noqa: E402
from oracle import host_ops as H, network as N  # noqa: E402

S = 256
B = 8
ORACLE_CFG = dict(PRE_NMS_LIMIT=6000, POST_NMS_ROIS_INFERENCE=1000, RPN_NMS_THRESHOLD=0.7,
                  RPN_BBOX_STD_DEV=(0.1, 0.1, 0.2, 0.2), BBOX_STD_DEV=(0.1, 0.1, 0.2, 0.2),
                  DETECTION_MIN_CONFIDENCE=0, DETECTION_NMS_THRESHOLD=0.3, DETECTION_MAX_INSTANCES=100,
                  POOL_SIZE=7, MASK_POOL_SIZE=14)

BOUNDS = dict(matched_frac=0.88, dbox_px_median=0.03, dbox_px_p95=0.08, dscore_median=6e-3, dscore_p95=2e-2,
              mask_iou_median=0.99, mask_iou_ge_09_frac=0.90)


@pytest.fixture(scope="module")
def weights():
    return N.make_random_weights(0, 4)


@pytest.fixture(scope="module")
def model(weights):
    from mrcnn import model as modellib
    from mrcnn.config import Config

    class C(Config):
        NAME = "rg-dataset"
        GPU_COUNT = 1
        IMAGES_PER_GPU = B
        NUM_CLASSES = 4
        IMAGE_MIN_DIM = S
        IMAGE_MAX_DIM = S
        RPN_ANCHOR_SCALES = (4, 8, 16, 32, 64)
        MEAN_PIXEL = np.array([0, 0, 0])
        DETECTION_MIN_CONFIDENCE = 0
        RPN_NMS_THRESHOLD = 0.7
    m = modellib.MaskRCNN(mode="inference", config=C(), model_dir="/tmp/mrcnn_logs")
    m.set_weights(weights)
    return m


def _oracle_detect(net, m, anchors):
    img = H.fits_to_rgb(m)
    molded, metas, windows = H.mold_inputs([img], min_dim=S, max_dim=S, min_scale=0, mode="square",
                                           mean_pixel=np.array([0, 0, 0]), num_classes=4)
    out = net.predict(molded, metas, anchors, ORACLE_CFG)
    b, c, sc, mk = H.unmold_detections(out["detections"][0], out["mrcnn_mask"][0], img.shape, (S, S, 3), windows[0])
    return out["detections"][0], {"rois": b, "class_ids": c, "scores": sc, "masks": mk}


def _check(sd, sr, what):
    msg = "%s: %r / %r" % (what, sd, sr)
    assert sd["matched_frac"] >= BOUNDS["matched_frac"], msg
    assert sd["dbox_px_median"] <= BOUNDS["dbox_px_median"] and sd["dbox_px_p95"] <= BOUNDS["dbox_px_p95"], msg
    assert sd["dscore_median"] <= BOUNDS["dscore_median"] and sd["dscore_p95"] <= BOUNDS["dscore_p95"], msg
    assert sr["mask_iou_median"] >= BOUNDS["mask_iou_median"], msg
    assert sr["mask_iou_ge_0.9_frac"] >= BOUNDS["mask_iou_ge_09_frac"], msg


def test_final_detections_vs_fp32_oracle_on_synthetic_maps(model, weights):
    """configs[1]: 16 of the 64 synthetic maps."""
    net = N.OracleNet(weights, 4, emulate_bf16=False)
    anchors = H.get_anchors((S, S, 3), (4, 8, 16, 32, 64))
    m_det, m_res = [], []
    for b0 in (0, B):
        maps = synth.radio_maps(B, S, start=b0)
        res = model.detect_maps(maps)
        det = model.read_tensor("detections")
        for i in range(B):
            odet, ores = _oracle_detect(net, maps[i], anchors)
            m_det.append(PM.match_detections_tensor(det[i], odet, S))
            m_res.append(PM.match_results(res[i], ores))
    sd, sr = PM.summarize(m_det), PM.summarize(m_res)
    print("e2e parity (16 synthetic maps, bf16 engine vs fp32 oracle):", sd, sr)
    out = os.environ.get("MRCNN_PARITY_JSON")
    if out:
        import json
        json.dump({"detections_tensor": sd, "unmolded": sr}, open(out, "w"), indent=1)
    _check(sd, sr, "synthetic maps")


def test_final_detections_vs_fp32_oracle_on_galaxy0002(model, weights, golden_dir):
    """configs[0]: the shipped 132x132 FITS map (resized x1.939 to 256, 288 NaN pixels)."""
    raw, _ = H.parse_fits_primary(open(os.path.join(golden_dir, "galaxy0002.fits"), "rb").read())
    m = np.ascontiguousarray(raw, dtype=np.float32).reshape(raw.shape[-2:])
    net = N.OracleNet(weights, 4, emulate_bf16=False)
    anchors = H.get_anchors((S, S, 3), (4, 8, 16, 32, 64))
    res = model.detect_maps(np.stack([m] * B))
    det = model.read_tensor("detections")
    assert all(np.array_equal(det[0], det[i]) for i in range(1, B)), "identical maps of one batch must give identical detections"
    odet, ores = _oracle_detect(net, m, anchors)
    sd = PM.summarize([PM.match_detections_tensor(det[0], odet, S)])
    sr = PM.summarize([PM.match_results(res[0], ores)])
    print("e2e parity (galaxy0002.fits):", sd, sr)
    # a single image: same bounds on the medians, looser on the tails
    assert sd["matched_frac"] >= 0.85 and sd["dbox_px_median"] <= BOUNDS["dbox_px_median"], (sd, sr)
    assert sd["dscore_median"] <= BOUNDS["dscore_median"] and sr["mask_iou_median"] >= BOUNDS["mask_iou_median"], (sd, sr)


def test_mask_bits_kernel_matches_byte_kernel():
    """mrcnn_unmold_detections_bits (what crosses PCIe) against mrcnn_unmold_detections ([B,H,W,D] bytes) through the
    C ABI: identical boxes / ids / scores / counts and bit-for-bit identical masks, for 1-, 2-, 4- and 8-word pixels."""
    import ctypes
    from mrcnn import _native
    lib = _native.lib()
    rng = np.random.default_rng(5)
    for D, (H0, W0), Bn in ((100, (100, 132), 3), (40, (64, 64), 2), (200, (48, 80), 2), (20, (33, 47), 1)):
        NC, MH = 4, 28
        det = np.zeros((Bn, D, 6), np.float32)
        for b in range(Bn):
            n = int(rng.integers(0, D + 1)) if b else D
            y1, x1 = rng.uniform(0, 0.8, (2, n))
            det[b, :n, 0], det[b, :n, 1] = y1, x1
            det[b, :n, 2] = np.minimum(1.0, y1 + rng.uniform(0.0, 0.5, n))      # some zero-area rows after rounding
            det[b, :n, 3] = np.minimum(1.0, x1 + rng.uniform(0.0, 0.5, n))
            det[b, :n, 4] = rng.integers(1, NC, n)
            det[b, :n, 5] = rng.uniform(0, 1, n)
        masks = rng.uniform(0, 1, (Bn, D, MH, MH, NC)).astype(np.float32)
        win = np.tile(np.array([[0, 0, 256, 256]], np.int32), (Bn, 1))
        d_det, d_m, d_win = (torch.from_numpy(a).cuda() for a in (det, masks, win))
        outs = []
        dw = lib.mrcnn_mask_bits_words(D)
        for bits in (False, True):
            rois = torch.zeros((Bn, D, 4), dtype=torch.int32, device="cuda")
            cls = torch.zeros((Bn, D), dtype=torch.int32, device="cuda")
            sc = torch.zeros((Bn, D), dtype=torch.float32, device="cuda")
            cnt = torch.zeros((Bn,), dtype=torch.int32, device="cuda")
            out = (torch.zeros((Bn, H0 * W0, dw), dtype=torch.int32, device="cuda") if bits
                   else torch.zeros((Bn, H0, W0, D), dtype=torch.uint8, device="cuda"))
            ws = torch.empty((lib.mrcnn_unmold_workspace_bytes(Bn, D),), dtype=torch.uint8, device="cuda")
            fn = lib.mrcnn_unmold_detections_bits if bits else lib.mrcnn_unmold_detections
            _native.check(fn(_native.ptr(d_det), _native.ptr(d_m), Bn, D, MH, MH, NC, (ctypes.c_int * 2)(H0, W0),
                             (ctypes.c_int * 2)(256, 256), _native.ptr(d_win), _native.ptr(rois), _native.ptr(cls), _native.ptr(sc),
                             _native.ptr(cnt), _native.ptr(out), _native.ptr(ws), ws.numel(), None))
            torch.cuda.synchronize()
            outs.append([t.cpu().numpy() for t in (rois, cls, sc, cnt, out)])
        (r0, c0, s0, n0, m0), (r1, c1, s1, n1, m1) = outs
        assert np.array_equal(r0, r1) and np.array_equal(c0, c1) and np.array_equal(s0, s1) and np.array_equal(n0, n1)
        assert n0.max() > 0
        unpacked = np.unpackbits(m1.view(np.uint8).reshape(Bn, H0 * W0, dw * 4), axis=2, bitorder="little")[:, :, :D]
        assert np.array_equal(unpacked.reshape(Bn, H0, W0, D), m0), (D, H0, W0)
        # and the host expansion of those bits = the reference's per-image [H,W,N] arrays
        dense = np.zeros((Bn, H0 * W0 * D), np.uint8)
        dst = (ctypes.c_void_p * Bn)(*[dense[i].ctypes.data for i in range(Bn)])
        bits_host = np.ascontiguousarray(m1.view(np.uint32))
        _native.check(lib.mrcnn_host_expand_mask_bits(bits_host.ctypes.data, Bn, H0 * W0, dw, n0.ctypes.data, dst, 0))
        for i in range(Bn):
            n = int(n0[i])
            assert np.array_equal(dense[i, :H0 * W0 * n].reshape(H0, W0, n), m0[i, :, :, :n])
