"""GPU parity of the train-mode kernels and graph (through the C ABI) against oracle/train_ops.py.
Bit-exact for the DetectionTargetLayer (indices, boxes, deltas, rounded masks); stated tolerances for gradients."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

import synth  # noqa: E402
from oracle import train_ops as TO  # noqa: E402


def _native():
    from mrcnn import _native
    return _native


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _blobs(rng, S, n):
    yy, xx = np.mgrid[0:S, 0:S]
    masks = np.zeros((S, S, n), dtype=bool)
    for i in range(n):
        cy, cx = rng.uniform(0.1 * S, 0.9 * S, 2)
        ry, rx = rng.uniform(3, 0.15 * S, 2)
        masks[:, :, i] = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0
    return masks


def _boxes_of(masks):
    from mrcnn import utils
    return utils.extract_bboxes(masks)


def _targets_case(rng, S, N, G, n_obj, crowd=False, T=64, mini=False):
    masks = _blobs(rng, S, n_obj)
    keep = masks.sum((0, 1)) > 0
    masks = masks[:, :, keep]
    n_obj = masks.shape[-1]
    boxes_px = _boxes_of(masks)
    cls = rng.integers(1, 4, n_obj).astype(np.int32)
    if crowd:
        cls[::3] *= -1
    gt_boxes = np.zeros((G, 4), np.float32)
    gt_cls = np.zeros((G,), np.int32)
    gt_masks = np.zeros((S, S, G), bool)
    gt_boxes[:n_obj] = ((boxes_px - np.array([0, 0, 1, 1])) / np.float32(S - 1)).astype(np.float32)
    gt_cls[:n_obj] = cls
    gt_masks[:, :, :n_obj] = masks
    if mini:
        from mrcnn import utils
        mm = utils.minimize_mask(boxes_px, masks, (56, 56))
        gt_masks = np.zeros((56, 56, G), bool)
        gt_masks[:, :, :n_obj] = mm
    # proposals: jittered GT boxes (positives), random boxes (negatives), zero rows in the middle and at the end
    props = np.zeros((N, 4), np.float32)
    k = 0
    for j in range(min(n_obj * 6, N // 2)):
        b = gt_boxes[j % n_obj] + rng.normal(0, 0.02, 4).astype(np.float32)
        props[k] = np.clip(b, 0, 1)
        k += 1
    nrand = (N - k) * 2 // 3
    yx = np.sort(rng.random((nrand, 2, 2)).astype(np.float32), axis=1)
    props[k:k + nrand] = np.stack([yx[:, 0, 0], yx[:, 0, 1], yx[:, 1, 0], yx[:, 1, 1]], 1)
    props[5] = 0
    return props, gt_cls, gt_boxes, gt_masks


def run_detection_targets(props, gt_cls, gt_boxes, gt_masks, T, ratio, mask_shape, mini, seed):
    nat = _native()
    lib = nat.lib()
    B, N = props.shape[:2]
    G = gt_cls.shape[1]
    d = [_dev(props), _dev(gt_cls), _dev(gt_boxes), _dev(gt_masks.astype(np.uint8))]
    rois = torch.empty((B, T, 4), dtype=torch.float32, device="cuda")
    tcls = torch.empty((B, T), dtype=torch.int32, device="cuda")
    tbox = torch.empty((B, T, 4), dtype=torch.float32, device="cuda")
    tmask = torch.empty((B, T) + tuple(mask_shape), dtype=torch.float32, device="cuda")
    counts = torch.empty((B, 2), dtype=torch.int32, device="cuda")
    sd = nat.float_array([0.1, 0.1, 0.2, 0.2])
    nat.check(lib.mrcnn_detection_targets(nat.ptr(d[0]), nat.ptr(d[1]), nat.ptr(d[2]), nat.ptr(d[3]), B, N, G, gt_masks.shape[1],
                                          gt_masks.shape[2], int(mini), T, ratio, sd, mask_shape[0], mask_shape[1], seed, None,
                                          nat.ptr(rois), nat.ptr(tcls), nat.ptr(tbox), nat.ptr(tmask), nat.ptr(counts), None),
              "detection_targets")
    torch.cuda.synchronize()
    return [t.cpu().numpy() for t in (rois, tcls, tbox, tmask, counts)]


def test_shuffle_key_host_export_equals_oracle():
    lib = _native().lib()
    for args in ((0, 0, 0, 0), (123456789012345, 3, 1, 1999), (2 ** 63 + 5, 63, 0, 77)):
        assert lib.mrcnn_shuffle_key(*args) == TO.shuffle_key(*args)


@pytest.mark.parametrize("case", ["plain", "crowd", "few_gt", "no_gt", "mini_mask"])
def test_detection_targets_bit_exact(case):
    rng = np.random.default_rng(["plain", "crowd", "few_gt", "no_gt", "mini_mask"].index(case))
    S, N, G, T = 128, 300, 16, 64
    n_obj = {"plain": 9, "crowd": 10, "few_gt": 1, "no_gt": 9, "mini_mask": 7}[case]
    batch = [_targets_case(rng, S, N, G, n_obj, crowd=(case == "crowd"), T=T, mini=(case == "mini_mask")) for _ in range(3)]
    if case == "no_gt":
        batch[1] = (batch[1][0], np.zeros_like(batch[1][1]), np.zeros_like(batch[1][2]), np.zeros_like(batch[1][3]))
    props, gt_cls, gt_boxes, gt_masks = [np.stack(x) for x in zip(*batch)]
    seed = 4242 + len(case)
    got = run_detection_targets(props, gt_cls, gt_boxes, gt_masks, T, 0.33, (28, 28), case == "mini_mask", seed)
    for b in range(props.shape[0]):
        rois, cls, deltas, masks, (pc, ncnt) = TO.detection_targets(props[b], gt_cls[b], gt_boxes[b], gt_masks[b], T, 0.33,
                                                                    (0.1, 0.1, 0.2, 0.2), (28, 28), case == "mini_mask", seed, b)
        assert got[4][b].tolist() == [pc, ncnt], (case, b)
        assert np.array_equal(got[0][b].view(np.uint32), rois.view(np.uint32)), "rois (image %d)" % b
        assert np.array_equal(got[1][b], cls)
        assert np.array_equal(got[2][b].view(np.uint32), deltas.view(np.uint32)), "deltas (image %d)" % b
        assert np.array_equal(got[3][b], masks), "mask targets (image %d)" % b
        if case in ("plain", "mini_mask"):
            assert pc > 0 and ncnt > 0 and masks[:pc].sum() > 0


def test_roi_align_backward_matches_autograd_of_the_oracle():
    nat = _native()
    lib = nat.lib()
    rng = np.random.default_rng(9)
    B, N, C, S, pool = 2, 24, 64, 128, 7
    shapes = [(32, 32), (16, 16), (8, 8), (4, 4)]
    feats = [rng.normal(0, 1, (B, h, w, C)).astype(np.float32) for h, w in shapes]
    yx = np.sort(rng.random((B, N, 2, 2)).astype(np.float32), axis=2)
    boxes = np.stack([yx[:, :, 0, 0], yx[:, :, 0, 1], yx[:, :, 1, 0], yx[:, :, 1, 1]], -1)
    boxes[0, 0] = [0.2, 0.3, 0.9, 1.2]           # partly outside the map
    dout = rng.normal(0, 1, (B, N, pool, pool, C)).astype(np.float32)
    d_feats = [_dev(f).to(torch.bfloat16) for f in feats]
    d_boxes = _dev(boxes)
    out = torch.empty((B, N, pool, pool, C), dtype=torch.bfloat16, device="cuda")
    levels = torch.empty((B, N), dtype=torch.int32, device="cuda")
    ptrs = (ctypes.c_void_p * 4)(*[f.data_ptr() for f in d_feats])
    hs = (ctypes.c_int * 4)(*[s[0] for s in shapes])
    ws = (ctypes.c_int * 4)(*[s[1] for s in shapes])
    nat.check(lib.mrcnn_pyramid_roi_align(ptrs, hs, ws, C, nat.DTYPE_BF16, nat.ptr(d_boxes), B, N, pool, float(S * S),
                                          nat.ptr(out), nat.ptr(levels), None), "roi_align")
    d_dout = _dev(dout).to(torch.bfloat16)
    grads = [torch.zeros((B, h, w, C), dtype=torch.float32, device="cuda") for h, w in shapes]
    gptrs = (ctypes.c_void_p * 4)(*[g.data_ptr() for g in grads])
    nat.check(lib.mrcnn_pyramid_roi_align_backward(gptrs, hs, ws, C, nat.ptr(d_boxes), nat.ptr(levels), B, N, pool,
                                                   nat.ptr(d_dout), None), "roi_align_backward")
    torch.cuda.synchronize()
    lv = levels.cpu().numpy()
    tf = [torch.tensor(f, requires_grad=True) for f in feats]
    g_bf = d_dout.float().cpu()
    total = 0
    for b in range(B):
        for n in range(N):
            crop = TO.crop_and_resize(tf[lv[b, n] - 2][b], boxes[b, n:n + 1], pool)[0]
            total = total + (crop * g_bf[b, n]).sum()
    total.backward()
    for l in range(4):
        want = tf[l].grad.numpy() if tf[l].grad is not None else np.zeros_like(feats[l])
        got = grads[l].cpu().numpy()
        assert np.abs(got - want).max() <= 1e-4 * max(1.0, np.abs(want).max()), "level %d" % (l + 2)


def test_sgd_step_matches_keras_restatement():
    nat = _native()
    lib = nat.lib()
    rng = np.random.default_rng(3)
    sizes = [("a", "kernel", 1000), ("a", "bias", 10), ("bn", "gamma", 10), ("b", "kernel", 77)]
    starts, off = [], 0
    for _, _, n in sizes:
        starts.append(off)
        off += (n + 7) // 8 * 8
    starts.append(off)
    w = np.zeros(off, np.float32)
    g = np.zeros(off, np.float32)
    v = np.zeros(off, np.float32)
    W, G, V = {}, {}, {}
    for (name, role, n), s in zip(sizes, starts):
        w[s:s + n] = rng.normal(0, 1, n)
        g[s:s + n] = rng.normal(0, 3, n)
        v[s:s + n] = rng.normal(0, 0.1, n)
        W[(name, role)], G[(name, role)], V[(name, role)] = w[s:s + n].copy(), g[s:s + n].copy(), v[s:s + n].copy()
    wd, lr, mom, clip, world = 1e-2, 0.05, 0.9, 5.0, 2
    coefs = np.array([2 * wd / n if role in ("kernel", "bias") else 0.0 for _, role, n in sizes], np.float32)
    dw, dg, dv = _dev(w), _dev(g), _dev(v)
    wb = torch.zeros(off, dtype=torch.bfloat16, device="cuda")
    sumsq = torch.zeros(1, dtype=torch.float64, device="cuda")
    d_starts, d_coefs = _dev(np.array(starts, np.int64)), _dev(coefs)
    nat.check(lib.mrcnn_sgd_step(nat.ptr(dg), nat.ptr(dw), nat.ptr(dv), nat.ptr(wb), off, nat.ptr(d_starts),
                                 nat.ptr(d_coefs), len(sizes), 1.0 / world, clip, lr, mom, nat.ptr(sumsq), None), "sgd_step")
    torch.cuda.synchronize()
    norm = TO.sgd_step(W, G, V, lr, mom, clip, wd, world=world)
    assert norm > clip                                # the clipping branch is exercised
    assert abs(float(sumsq.sqrt()) - norm) <= 1e-5 * norm
    gw, gv = dw.cpu().numpy(), dv.cpu().numpy()
    for (name, role, n), s in zip(sizes, starts):
        assert np.allclose(gw[s:s + n], W[(name, role)], rtol=0, atol=2e-6), (name, role)
        assert np.allclose(gv[s:s + n], V[(name, role)], rtol=0, atol=2e-6), (name, role)
    assert torch.equal(wb, dw.to(torch.bfloat16))


@pytest.mark.parametrize("case", ["1x1_small", "1x1_fc1", "3x3_mask", "3x3_backbone", "1x1_cout_ragged"])
def test_conv_wgrad_tcgen05_vs_torch(case):
    """mrcnn_conv2d_wgrad_bf16 (MN-major tcgen05 GEMM, split K, float32 atomics) against torch's float32 weight gradient
    of the same bf16 tensors; dw is accumulated on top of what the buffer held."""
    nat = _native()
    lib = nat.lib()
    n, h, w, cin, cout, k = {"1x1_small": (2, 16, 16, 64, 128, 1), "1x1_fc1": (1, 25, 8, 1024, 256, 1),
                             "3x3_mask": (37, 14, 14, 256, 256, 3), "3x3_backbone": (2, 32, 32, 128, 128, 3),
                             "1x1_cout_ragged": (3, 9, 11, 192, 72, 1)}[case]
    rng = np.random.default_rng(len(case))
    x = torch.from_numpy(rng.normal(0, 1, (n, h, w, cin)).astype(np.float32)).cuda().to(torch.bfloat16)
    dy = torch.from_numpy(rng.normal(0, 1, (n, h, w, cout)).astype(np.float32)).cuda().to(torch.bfloat16)
    base = torch.from_numpy(rng.normal(0, 1, (cout, k, k, cin)).astype(np.float32)).cuda()
    dw = base.clone()
    desc = nat.ConvDesc(n=n, h=h, w=w, cin=cin, kh=k, kw=k, stride=1, pad=k // 2, cout=cout, relu=0, residual_upsample2=0,
                        out_dtype=nat.DTYPE_F32, out_mode=0, out_ld=0)
    wts = torch.from_numpy(rng.normal(0, 1, (cout, k, k, cin)).astype(np.float32)).cuda().to(torch.bfloat16)
    wdot = torch.zeros(cout, dtype=torch.float32, device="cuda")
    nat.check(lib.mrcnn_conv2d_wgrad_bf16(ctypes.byref(desc), nat.ptr(x), nat.ptr(dy), nat.ptr(dw), nat.ptr(wts), nat.ptr(wdot), None),
              "conv2d_wgrad")
    torch.cuda.synchronize()
    want = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (cout, cin, k, k), dy.float().permute(0, 3, 1, 2),
                                       stride=1, padding=k // 2).permute(0, 2, 3, 1)
    got = (dw - base).cpu().numpy()
    want = want.cpu().numpy()
    err = np.abs(got - want).max()
    assert err <= 2e-3 * np.abs(want).max() + 1e-3, (case, err, np.abs(want).max())
    want_dot = (wts.float().cpu().numpy() * want).sum((1, 2, 3))
    assert np.abs(wdot.cpu().numpy() - want_dot).max() <= 2e-3 * np.abs(want_dot).max() + 1e-2, case


@pytest.mark.parametrize("case", ["1x1", "1x1_wide", "3x3_mask", "3x3_backbone", "3x3_cin64"])
def test_conv_dgrad_tcgen05_vs_torch(case):
    """mrcnn_conv2d_dgrad_bf16 (the forward implicit GEMM with its B operand read MN-major from the forward weights,
    filter flipped by coordinates) against torch's float32 conv2d_input on the same bf16 tensors."""
    nat = _native()
    lib = nat.lib()
    n, h, w, cin, cout, k = {"1x1": (2, 16, 16, 256, 64, 1), "1x1_wide": (1, 32, 32, 1024, 256, 1), "3x3_mask": (19, 14, 14, 256, 256, 3),
                             "3x3_backbone": (2, 32, 32, 128, 128, 3), "3x3_cin64": (3, 17, 23, 64, 192, 3)}[case]
    rng = np.random.default_rng(len(case))
    dy = torch.from_numpy(rng.normal(0, 1, (n, h, w, cout)).astype(np.float32)).cuda().to(torch.bfloat16)
    wts = torch.from_numpy(rng.normal(0, 0.05, (cout, k, k, cin)).astype(np.float32)).cuda().to(torch.bfloat16)
    dx = torch.empty((n, h, w, cin), dtype=torch.bfloat16, device="cuda")
    ones = torch.ones(cin, dtype=torch.float32, device="cuda")
    zeros = torch.zeros(cin, dtype=torch.float32, device="cuda")
    desc = nat.ConvDesc(n=n, h=h, w=w, cin=cin, kh=k, kw=k, stride=1, pad=k // 2, cout=cout, relu=0, residual_upsample2=0,
                        out_dtype=nat.DTYPE_BF16, out_mode=0, out_ld=0)
    nat.check(lib.mrcnn_conv2d_dgrad_bf16(ctypes.byref(desc), nat.ptr(dy), nat.ptr(wts), nat.ptr(ones), nat.ptr(zeros), nat.ptr(dx), None),
              "conv2d_dgrad")
    torch.cuda.synchronize()
    want = torch.nn.grad.conv2d_input((n, cin, h, w), wts.float().permute(0, 3, 1, 2), dy.float().permute(0, 3, 1, 2), stride=1,
                                      padding=k // 2).permute(0, 2, 3, 1).cpu().numpy()
    got = dx.float().cpu().numpy()
    assert np.abs(got - want).max() <= 1e-2 * np.abs(want).max() + 1e-3, (case, np.abs(got - want).max(), np.abs(want).max())


# ---------------------------------------------------------------------------------------------------------------
# the whole training graph on a tiny configuration
# ---------------------------------------------------------------------------------------------------------------

def _tiny_config():
    from mrcnn.config import Config

    class C(Config):
        NAME = "tiny_train"
        NUM_CLASSES = 4
        GPU_COUNT = 1
        IMAGES_PER_GPU = 1
        IMAGE_MIN_DIM = 128
        IMAGE_MAX_DIM = 128
        RPN_ANCHOR_SCALES = (4, 8, 16, 32, 64)
        MEAN_PIXEL = np.array([0, 0, 0])
        RPN_TRAIN_ANCHORS_PER_IMAGE = 64
        MAX_GT_INSTANCES = 12
        TRAIN_ROIS_PER_IMAGE = 32
        POST_NMS_ROIS_TRAINING = 200
        USE_MINI_MASK = False
    return C()


def _tiny_inputs(cfg, seed=0, gt_boxes_px=None):
    """gt_boxes_px: rectangles to use as ground truth (filled-rectangle masks) instead of random blobs — the tests pass a
    few of the graph's own proposals, which guarantees positive ROIs whatever the random weights propose."""
    from mrcnn import model as modellib, utils
    rng = np.random.default_rng(seed)
    S = 128
    if gt_boxes_px is None:
        masks = _blobs(rng, S, 6)
    else:
        masks = np.zeros((S, S, len(gt_boxes_px)), bool)
        for i, (y1, x1, y2, x2) in enumerate(gt_boxes_px):
            masks[y1:y2, x1:x2, i] = True
    masks = masks[:, :, masks.sum((0, 1)) > 0]
    cls = rng.integers(1, 4, masks.shape[-1]).astype(np.int32)
    boxes = utils.extract_bboxes(masks)
    img = (rng.random((S, S, 3)) * 40).astype(np.float32)
    for i in range(masks.shape[-1]):
        img[masks[:, :, i]] += 150
    anchors = utils.generate_pyramid_anchors(cfg.RPN_ANCHOR_SCALES, cfg.RPN_ANCHOR_RATIOS,
                                             utils.compute_backbone_shapes(cfg, cfg.IMAGE_SHAPE), cfg.BACKBONE_STRIDES, 1)
    np.random.seed(5)
    match, bbox = modellib.build_rpn_targets((S, S, 3), anchors, cls, boxes, cfg)
    G = cfg.MAX_GT_INSTANCES
    gcls, gbox, gm = np.zeros((1, G), np.int32), np.zeros((1, G, 4), np.int32), np.zeros((1, S, S, G), bool)
    n = len(cls)
    gcls[0, :n], gbox[0, :n], gm[0, :, :, :n] = cls, boxes, masks
    meta = modellib.compose_image_meta(0, (S, S, 3), (S, S, 3), (0, 0, S, S), 1.0, np.ones(4, np.int32))[None]
    return [img[None], meta, match[None, :, None], bbox[None], gcls, gbox, gm]


def _inputs_with_positive_rois(g, cfg):
    with torch.no_grad():
        g.forward(g.to_device(_tiny_inputs(cfg)), seed=1)
    rois = g.taps["rpn_rois"][0].cpu().numpy()
    px = np.round(rois * 127.0 + np.array([0, 0, 1, 1])).astype(int)
    px = np.clip(px, 0, 128)
    ok = [tuple(b) for b in px if b[2] - b[0] >= 3 and b[3] - b[1] >= 3]
    ok = sorted(set(ok), key=lambda b: -(b[2] - b[0]) * (b[3] - b[1]))       # the largest distinct proposals
    assert len(ok) >= 3, "the random network proposed no usable box: %s" % px[:8].tolist()
    return _tiny_inputs(cfg, gt_boxes_px=ok[:5])


def test_training_graph_losses_and_gradients_vs_fp32_oracle():
    """bf16 graph on the GPU vs the fp32 torch-CPU oracle fed with the product's own ROIs and targets (chain of custody:
    proposals and targets are index stages checked bit-exactly elsewhere).  Tolerances: every loss within 3 % + 0.02 abs,
    gradient cosine similarity >= 0.97 per checked layer (bf16 activations through 101 layers vs fp32)."""
    from mrcnn import training
    cfg = _tiny_config()
    weights = synth.make_random_weights(0, 4)
    g = training.TrainGraph(cfg, layers="all", seed=11)
    g.params.set_weights(weights)
    inputs = _inputs_with_positive_rois(g, cfg)
    dev = g.to_device(inputs)
    g.params.g.zero_()
    total, ls = g.forward(dev, seed=11)
    total.backward()
    g.finish_backward()
    torch.cuda.synchronize()
    taps = {k: v.detach().cpu() for k, v in g.taps.items()}
    assert int(taps["counts"][0, 0]) > 0, "no positive ROI in the tiny case"
    # the graph's own DetectionTargetLayer call (GT boxes normalised inside the graph) against the oracle, bit for bit
    gt_norm = ((inputs[5][0].astype(np.float32) - np.array([0, 0, 1, 1], np.float32)) / np.float32(127)).astype(np.float32)
    o = TO.detection_targets(taps["rpn_rois"][0].numpy(), inputs[4][0], gt_norm, inputs[6][0], cfg.TRAIN_ROIS_PER_IMAGE, cfg.ROI_POSITIVE_RATIO,
                             (0.1, 0.1, 0.2, 0.2), (28, 28), False, 11, 0)
    assert np.array_equal(taps["rois"][0].numpy().view(np.uint32), o[0].view(np.uint32))
    assert np.array_equal(taps["target_class_ids"][0].numpy(), o[1])
    assert np.array_equal(taps["target_bbox"][0].numpy().view(np.uint32), o[2].view(np.uint32))
    assert np.array_equal(taps["target_mask"][0].numpy(), o[3])
    # the oracle on the same ROIs / targets
    net = TO.TrainNet(weights, cfg)
    P = net.backbone_fpn(inputs[0])
    lg, _, rb = net.rpn(P)
    rois = taps["rois"].numpy()
    logits, bbox = net.class_head(rois, P)
    masks = net.mask_head(rois, P)
    ols = TO.losses(torch.tensor(inputs[2]), torch.tensor(inputs[3]), lg, rb, taps["target_class_ids"], taps["target_bbox"],
                    taps["target_mask"], logits, bbox, masks, torch.ones((1, 4), dtype=torch.int32))
    ototal = sum(ols.values())
    ototal.backward()
    for k in training.LOSS_NAMES:
        a, b = float(ls[k].detach()), float(ols[k].detach())
        assert abs(a - b) <= 0.03 * abs(b) + 0.02, (k, a, b)
    checked = 0
    for (name, role), t in g.masters.items():
        if role != "kernel" or not (name.startswith(("mrcnn_", "rpn_", "fpn_")) or name in ("res5c_branch2c", "res4w_branch2b", "res2a_branch2a")):
            continue
        got = g.params.view(g.params.g, name, role).cpu().numpy()
        kind = g.params.kinds[name]
        want = net.p[(name, "kernel")].grad.numpy()
        want = {"conv": lambda a: a.transpose(3, 0, 1, 2), "dense": lambda a: a.T, "deconv": lambda a: a}[kind](want)
        if np.linalg.norm(want) == 0.0:             # e.g. fpn_p5 at 128x128: no ROI on level 5, no RPN sample on P5 / P6
            assert np.linalg.norm(got) == 0.0, name
            continue
        cos = float((got * want).sum() / (np.linalg.norm(got) * np.linalg.norm(want) + 1e-30))
        assert cos >= 0.97, (name, cos)
        ratio = float(np.linalg.norm(got) / np.linalg.norm(want))
        assert 0.9 <= ratio <= 1.1, (name, ratio)
        checked += 1
    assert checked >= 18
    # BatchNorm gamma / beta and bias gradients (batched affine backward of the fused layers)
    for name, role in (("bn4f_branch2b", "gamma"), ("bn4f_branch2b", "beta"), ("res4f_branch2b", "bias"), ("mrcnn_mask_bn2", "gamma"),
                       ("mrcnn_class_bn1", "beta"), ("fpn_p3", "bias"), ("rpn_conv_shared", "bias"), ("bn_conv1", "gamma")):
        got = g.params.view(g.params.g, name, role).cpu().numpy()
        want = net.p[(name, role)].grad.numpy()
        cos = float((got * want).sum() / (np.linalg.norm(got) * np.linalg.norm(want) + 1e-30))
        assert cos >= 0.95, (name, role, cos)
        assert 0.85 <= np.linalg.norm(got) / np.linalg.norm(want) <= 1.15, (name, role)


def test_train_steps_reduce_the_loss_on_a_fixed_batch():
    from mrcnn import training
    cfg = _tiny_config()
    g = training.TrainGraph(cfg, layers="all", seed=3)
    g.params.set_weights(synth.make_random_weights(0, 4))
    tr = training.Trainer(g, learning_rate=0.002, momentum=0.9)
    dev = g.to_device(_inputs_with_positive_rois(g, cfg))
    first = last = None
    for i in range(12):
        ls = tr.train_step(dev, seed=3)
        v = float(ls["loss"])
        assert np.isfinite(v)
        first = v if first is None else first
        last = v
    assert last < first, (first, last)
    assert tr.opt.grad_norm() > 0
    # the bf16 operand copy follows the master weights
    assert torch.equal(g.params.wb, g.params.w.to(torch.bfloat16))


def test_run_py_train_writes_a_checkpoint_the_detector_loads(tmp_path):
    """`run.py train` end to end on a tiny FITS dataset (reference scripts/run.py train -> MaskRCNN.train): one epoch of two
    steps + one validation step, a Keras-layout weights file per epoch, and that file loads into the inference engine."""
    import glob
    import importlib.util
    import os
    from mrcnn import fitsio
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "caesar-mrcnn_b200", "scripts", "run.py")
    spec = importlib.util.spec_from_file_location("run_b200_train_gpu", path)
    run = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(run)
    rng = np.random.default_rng(2)
    lines = []
    for i in range(6):
        m, masks, cls = synth.training_sample(i, 128)
        ipath = str(tmp_path / ("img%d.fits" % i))
        fitsio.write_primary(ipath, m)
        k = int(np.argmax(masks.sum((0, 1))))
        mpath = str(tmp_path / ("mask%d.fits" % i))
        fitsio.write_primary(mpath, masks[:, :, k].astype(np.float32))
        lines.append("%s,%s,%s" % (ipath, mpath, ["sidelobe", "source", "galaxy"][int(cls[k]) - 1]))
    lst = str(tmp_path / "list.dat")
    open(lst, "w").write("\n".join(lines) + "\n")
    logs = str(tmp_path / "logs")
    cwd = os.getcwd()
    os.chdir(str(tmp_path))
    try:
        rc = run.main(["train", "--datalist", lst, "--imgsize", "128", "--nimg_per_gpu", "1", "--nepochs", "1", "--epoch_length", "2",
                       "--nvalidation_steps", "1", "--validation_data_fract", "0.34", "--logs", logs,
                       "--train_rois_per_image", "32", "--rpn_train_anchors_per_image", "64", "--max_gt_instances", "8"])
    finally:
        os.chdir(cwd)
    assert rc == 0
    files = glob.glob(os.path.join(logs, "*", "mask_rcnn_rg-dataset_0001.h5"))
    assert len(files) == 1, os.listdir(logs)
    rc = run.main(["detect", "--image", lines[0].split(",")[0], "--weights", files[0], "--imgsize", "128",
                   "--detect_outfile_json", str(tmp_path / "det.json"), "--logs", logs])
    assert rc == 0
