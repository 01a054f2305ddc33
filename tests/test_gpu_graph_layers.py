"""GPU parity (through the C ABI): ProposalLayer, PyramidROIAlign, DetectionLayer vs the oracle.
Bit-exact for indices / levels / float32 samples; see oracle/graph_layers.py."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import graph_layers as GL, host_ops as H  # noqa: E402


def _native():
    from mrcnn import _native
    return _native


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def run_proposal(rpn_class, rpn_bbox, anchors, K=6000, R=1000, thr=0.7):
    nat = _native()
    lib = nat.lib()
    B, A = rpn_class.shape[:2]
    Kc = min(K, A)
    d_cls, d_box, d_anc = _dev(rpn_class), _dev(rpn_bbox), _dev(anchors)
    rois = torch.empty((B, R, 4), dtype=torch.float32, device="cuda")
    topk = torch.empty((B, Kc), dtype=torch.int32, device="cuda")
    keep = torch.empty((B, R), dtype=torch.int32, device="cuda")
    cnt = torch.empty((B,), dtype=torch.int32, device="cuda")
    sd = nat.float_array([0.1, 0.1, 0.2, 0.2])
    st = lib.mrcnn_proposal_layer(nat.ptr(d_cls), nat.ptr(d_box), nat.ptr(d_anc), int(anchors.ndim == 3), B, A, K, R,
                                  thr, sd, nat.ptr(rois), nat.ptr(topk), nat.ptr(keep), nat.ptr(cnt), None, 0, None)
    nat.check(st, "proposal_layer")
    torch.cuda.synchronize()
    return rois.cpu().numpy(), topk.cpu().numpy(), keep.cpu().numpy(), cnt.cpu().numpy()


def _rpn_inputs(rng, B, A, S=256, ties=False, sharp=2.0):
    anchors = H.get_anchors((S, S, 3), (4, 8, 16, 32, 64))[:A]
    logits = rng.normal(0, sharp, size=(B, A, 2)).astype(np.float32)
    e = np.exp(logits - logits.max(-1, keepdims=True))
    probs = (e / e.sum(-1, keepdims=True)).astype(np.float32)
    if ties:
        probs[:, ::3, 1] = np.float32(1.0)
        probs[:, 1::7, 1] = np.float32(0.5)
    bbox = rng.normal(0, 1.0, size=(B, A, 4)).astype(np.float32)
    return probs, bbox, anchors


def _check_proposal(probs, bbox, anchors, K=6000, R=1000, thr=0.7):
    rois, topk, keep, cnt = run_proposal(probs, bbox, anchors, K, R, thr)
    ref, taps = GL.proposal_layer(probs, bbox, anchors, pre_nms_limit=K, proposal_count=R, nms_threshold=thr,
                                  return_taps=True)
    for b in range(probs.shape[0]):
        assert np.array_equal(topk[b], taps[b]["topk"]), "top-k indices differ (image %d)" % b
        n = taps[b]["keep"].shape[0]
        assert cnt[b] == n
        assert np.array_equal(keep[b, :n], taps[b]["keep"]), "NMS keep indices differ (image %d)" % b
        assert np.all(keep[b, n:] == -1)
    assert np.array_equal(rois.view(np.uint32), ref.view(np.uint32)), "rois not bit-exact"


def test_proposal_reference_config_bit_exact():
    rng = np.random.default_rng(0)
    probs, bbox, anchors = _rpn_inputs(rng, 3, 16368)
    _check_proposal(probs, bbox, anchors)


def test_proposal_with_exact_score_ties():
    rng = np.random.default_rng(1)
    probs, bbox, anchors = _rpn_inputs(rng, 2, 16368, ties=True)
    _check_proposal(probs, bbox, anchors)


def test_proposal_fewer_anchors_than_limit_and_small_outputs():
    rng = np.random.default_rng(2)
    probs, bbox, anchors = _rpn_inputs(rng, 2, 1000)        # A < PRE_NMS_LIMIT
    _check_proposal(probs, bbox, anchors)
    probs, bbox, anchors = _rpn_inputs(rng, 1, 4092)
    _check_proposal(probs, bbox, anchors, K=600, R=50, thr=0.5)
    probs, bbox, anchors = _rpn_inputs(rng, 1, 77)           # ragged, tiny
    _check_proposal(probs, bbox, anchors, K=6000, R=1000)


def test_proposal_all_ties_and_zero_area():
    rng = np.random.default_rng(3)
    probs, bbox, anchors = _rpn_inputs(rng, 1, 8184)
    probs[..., 1] = np.float32(1.0)                           # every score identical
    _check_proposal(probs, bbox, anchors)
    anchors0 = anchors.copy()
    anchors0[:, 2] = anchors0[:, 0]                           # zero-height anchors -> zero-area boxes
    probs, bbox, _ = _rpn_inputs(rng, 1, 8184)
    _check_proposal(probs, bbox, anchors0)


TIE_KINDS = ["sparse_pairs", "many_pairs", "few_values", "quantised", "saturated_head", "triples"]


@pytest.mark.parametrize("kind", TIE_KINDS)
def test_proposal_tie_patterns_consumed_to_the_end(kind):
    """Differential test of the pipelined heap popper: tie patterns of every kind, and a low NMS threshold on heavily
    overlapping boxes so the NMS pops (almost) all K candidates instead of stopping at 1000 kept after ~1100 pops.
    The oracle's pop order is libstdc++'s std::priority_queue (oracle/nms_ref.cpp)."""
    rng = np.random.default_rng(40 + TIE_KINDS.index(kind))
    for A, K, R, thr in ((8184, 6000, 1000, 0.2), (16368, 6000, 1000, 0.05), (5000, 4097, 300, 0.3), (700, 6000, 1000, 0.1)):
        probs, bbox, anchors = _rpn_inputs(rng, 2, A)
        if thr < 0.25:                                          # big random boxes: nearly every pair overlaps
            cy, cx = rng.random((2, A)).astype(np.float32)
            hh, ww = (rng.random((2, A)) * 0.3 + 0.15).astype(np.float32)
            anchors = np.stack([cy - hh, cx - ww, cy + hh, cx + ww], 1).astype(np.float32)
        fg = rng.random((2, A)).astype(np.float32) * np.float32(0.98) + np.float32(0.01)
        if kind == "sparse_pairs":
            for b in range(2):
                idx = rng.integers(0, A - 1, 12)
                fg[b, idx] = fg[b, idx + 1]
        elif kind == "many_pairs":
            fg = (rng.integers(1, A // 3, (2, A)) / np.float32(A // 3 + 1)).astype(np.float32)
        elif kind == "few_values":
            fg = (rng.integers(1, 6, (2, A)) / np.float32(8)).astype(np.float32)
        elif kind == "quantised":
            fg = np.round(fg, 2).astype(np.float32) + np.float32(0.001)
        elif kind == "saturated_head":
            fg[:, rng.integers(0, A, A // 4)] = np.float32(1.0)
        elif kind == "triples":
            for b in range(2):
                idx = rng.integers(0, A - 2, 40)
                fg[b, idx] = fg[b, idx + 1] = fg[b, idx + 2]
        probs[..., 1] = fg
        probs[..., 0] = np.float32(1.0) - fg
        bbox *= np.float32(0.3)                                # boxes stay near their anchors: dense overlaps
        _check_proposal(probs, bbox, anchors, K=K, R=R, thr=thr)


def test_proposal_1024_config_261888_anchors():
    """Base-Config size (IMAGE_MAX_DIM = 1024): A = 261 888 anchors per image, top-6000 of them, with ties."""
    rng = np.random.default_rng(5)
    probs, bbox, anchors = _rpn_inputs(rng, 2, 261888, S=1024, sharp=3.0)
    probs[1, 5::4001, 1] = probs[1, 7, 1]                     # a few exact ties among the top scores
    _check_proposal(probs, bbox, anchors)


def test_proposal_batched_anchors():
    rng = np.random.default_rng(4)
    probs, bbox, anchors = _rpn_inputs(rng, 2, 4092)
    banch = np.broadcast_to(anchors, (2,) + anchors.shape).copy()
    banch[1] += np.float32(0.01)
    _check_proposal(probs, bbox, banch)


# ---------------------------------------------------------------------------------------------

def run_roialign(fmaps, boxes, pool, image_area, dtype="f32"):
    nat = _native()
    lib = nat.lib()
    B, N = boxes.shape[:2]
    C = fmaps[0].shape[-1]
    tdt = torch.float32 if dtype == "f32" else torch.bfloat16
    d_f = [_dev(f).to(tdt).contiguous() for f in fmaps]
    ptrs = (ctypes.c_void_p * 4)(*[f.data_ptr() for f in d_f])
    hs = (ctypes.c_int * 4)(*[f.shape[1] for f in fmaps])
    ws = (ctypes.c_int * 4)(*[f.shape[2] for f in fmaps])
    d_boxes = _dev(boxes)
    out = torch.empty((B, N, pool, pool, C), dtype=tdt, device="cuda")
    lv = torch.empty((B, N), dtype=torch.int32, device="cuda")
    st = lib.mrcnn_pyramid_roi_align(ptrs, hs, ws, C, nat.DTYPE_F32 if dtype == "f32" else nat.DTYPE_BF16,
                                     nat.ptr(d_boxes), B, N, pool, float(image_area), nat.ptr(out), nat.ptr(lv), None)
    nat.check(st, "pyramid_roi_align")
    torch.cuda.synchronize()
    return out.float().cpu().numpy(), lv.cpu().numpy()


def _pyramid(rng, B, S, C):
    return [rng.normal(0, 1, size=(B, S // s, S // s, C)).astype(np.float32) for s in (4, 8, 16, 32)]


def _mixed_boxes(rng, B, N):
    yx = rng.random((B, N, 2)).astype(np.float32) * 0.7
    scale = (2.0 ** rng.uniform(-6, 0, size=(B, N, 1))).astype(np.float32)
    hw = rng.random((B, N, 2)).astype(np.float32) * scale + np.float32(1e-3)
    boxes = np.concatenate([yx, np.minimum(yx + hw, 1.0)], axis=-1).astype(np.float32)
    boxes[:, :4] = 0.0                                        # zero-padded proposals
    boxes[:, 4] = [0.2, 0.2, 0.2, 0.9]                        # zero height
    boxes[:, 5] = [0.0, 0.0, 1.0, 1.0]                        # whole image
    boxes[:, 6] = [-0.1, 0.3, 1.2, 0.8]                       # partly outside -> zeros
    return boxes


@pytest.mark.parametrize("pool", [7, 14])
def test_roialign_fp32_bit_exact_and_levels(pool):
    rng = np.random.default_rng(10 + pool)
    B, S, C, N = 2, 256, 64, 300
    fm = _pyramid(rng, B, S, C)
    boxes = _mixed_boxes(rng, B, N)
    out, lv = run_roialign(fm, boxes, pool, S * S, "f32")
    ref, rlv = GL.pyramid_roi_align(boxes, (S, S, 3), fm, (pool, pool), return_levels=True)
    assert np.array_equal(lv, rlv), "ROI levels differ"
    assert len(np.unique(rlv)) >= 3                          # several pyramid levels exercised
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32)), "fp32 ROIAlign not bit-exact"


@pytest.mark.parametrize("pool,C", [(7, 256), (14, 256), (5, 256), (7, 64)])
def test_roialign_bf16_within_one_rounding(pool, C):
    """pool 7 / 14 with 256 channels = the engine's case (row-per-warp kernel); the others take the generic kernel"""
    rng = np.random.default_rng(20 + pool + C)
    B, S, N = 2, 128, 100
    fm = [torch.from_numpy(f).to(torch.bfloat16).float().numpy() for f in _pyramid(rng, B, S, C)]
    boxes = _mixed_boxes(rng, B, N)
    out, lv = run_roialign(fm, boxes, pool, S * S, "bf16")
    ref, rlv = GL.pyramid_roi_align(boxes, (S, S, 3), fm, (pool, pool), return_levels=True)
    assert np.array_equal(lv, rlv)
    ref_bf = torch.from_numpy(ref).to(torch.bfloat16).float().numpy()
    assert np.array_equal(out, ref_bf), "bf16 ROIAlign != bf16(round(fp32 oracle))"


def test_roi_levels_kernel_1024_config():
    nat = _native()
    rng = np.random.default_rng(30)
    boxes = _mixed_boxes(rng, 1, 5000)[0]
    d = _dev(boxes)
    lv = torch.empty((5000,), dtype=torch.int32, device="cuda")
    nat.check(nat.lib().mrcnn_roi_levels(nat.ptr(d), 5000, float(1024 * 1024), nat.ptr(lv), None))
    assert np.array_equal(lv.cpu().numpy(), GL.roi_levels(boxes, np.float32(1024 * 1024)))


# ---------------------------------------------------------------------------------------------

def run_detection(rois, probs, deltas, metas, D=100, min_conf=0.0, thr=0.3):
    nat = _native()
    lib = nat.lib()
    B, N, NC = probs.shape
    d = [_dev(x.astype(np.float32)) for x in (rois, probs, deltas, metas)]
    out = torch.empty((B, D, 6), dtype=torch.float32, device="cuda")
    sd = nat.float_array([0.1, 0.1, 0.2, 0.2])
    st = lib.mrcnn_detection_layer(nat.ptr(d[0]), nat.ptr(d[1]), nat.ptr(d[2]), nat.ptr(d[3]), metas.shape[1], B, N, NC,
                                   D, min_conf, thr, sd, nat.ptr(out), None)
    nat.check(st, "detection_layer")
    torch.cuda.synchronize()
    return out.cpu().numpy()


def _det_inputs(rng, B, N, NC, S=256, ties=False, window=(0, 0, 256, 256)):
    yx = rng.random((B, N, 2)).astype(np.float32) * 0.8
    hw = rng.random((B, N, 2)).astype(np.float32) * 0.25 + np.float32(0.01)
    rois = np.concatenate([yx, np.minimum(yx + hw, 1.0)], axis=-1).astype(np.float32)
    rois[:, N - 50:] = 0.0                                   # zero-padded proposals flow through
    logits = rng.normal(0, 2.0, size=(B, N, NC)).astype(np.float32)
    e = np.exp(logits - logits.max(-1, keepdims=True))
    probs = (e / e.sum(-1, keepdims=True)).astype(np.float32)
    if ties:
        probs[:, ::4] = np.array([0.1, 0.7, 0.1, 0.1][:NC], dtype=np.float32)
        probs[:, 1::9] = np.array([0.25, 0.25, 0.25, 0.25][:NC], dtype=np.float32)   # argmax tie -> class 0
    deltas = rng.normal(0, 1.0, size=(B, N, NC, 4)).astype(np.float32)
    metas = np.zeros((B, 12 + NC), dtype=np.float32)
    metas[:, 1:4] = [132, 132, 3]
    metas[:, 4:7] = [S, S, 3]
    metas[:, 7:11] = window
    metas[:, 11] = 1.9393939
    return rois, probs, deltas, metas


@pytest.mark.parametrize("ties", [False, True])
def test_detection_layer_bit_exact(ties):
    rng = np.random.default_rng(40 + int(ties))
    rois, probs, deltas, metas = _det_inputs(rng, 3, 1000, 4, ties=ties)
    out = run_detection(rois, probs, deltas, metas)
    ref = GL.detection_layer(rois, probs, deltas, metas, min_confidence=0.0)
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))
    assert (ref[..., 4] > 0).sum() > 50


def test_detection_layer_confidence_branch_window_and_few_rois():
    rng = np.random.default_rng(50)
    rois, probs, deltas, metas = _det_inputs(rng, 2, 1000, 4, window=(28, 0, 228, 256))
    out = run_detection(rois, probs, deltas, metas, min_conf=0.7)
    ref = GL.detection_layer(rois, probs, deltas, metas, min_confidence=0.7)
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))
    rois, probs, deltas, metas = _det_inputs(rng, 1, 60, 2)
    out = run_detection(rois, probs, deltas, metas, D=100)
    ref = GL.detection_layer(rois, probs, deltas, metas, min_confidence=0.0)
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))
    # all background -> all-zero output
    probs0 = np.zeros_like(probs)
    probs0[..., 0] = 1.0
    assert not run_detection(rois, probs0, deltas, metas).any()
