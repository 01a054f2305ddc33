"""CPU: independent cross-checks of the third-party semantics the oracle restates
(tf NMS vs torchvision, skimage<=0.15 resize vs cv2 interior pixels, crop_and_resize loop vs
vectorised twin, zscale sanity on the shipped FITS files)."""
import os

import numpy as np
import pytest
import torch

from oracle import _native, graph_layers as GL, host_ops as H


def _rand_boxes(rng, n, size=1.0):
    yx = rng.random((n, 2)).astype(np.float32) * 0.8 * size
    hw = (rng.random((n, 2)).astype(np.float32) * 0.2 + 0.01) * size
    return np.concatenate([yx, yx + hw], axis=1).astype(np.float32)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_nms_matches_torchvision_without_ties(seed):
    import torchvision
    rng = np.random.default_rng(seed)
    b = _rand_boxes(rng, 800)
    s = rng.permutation(800).astype(np.float32) / 800.0            # distinct scores
    keep = _native.nms_tf113(b, s, 800, 0.5)
    tv = torchvision.ops.nms(torch.from_numpy(b[:, [1, 0, 3, 2]].copy()), torch.from_numpy(s), 0.5).numpy()
    assert np.array_equal(keep, tv)
    # truncation at max_out
    assert np.array_equal(_native.nms_tf113(b, s, 17, 0.5), tv[:17])


def test_nms_properties_and_zero_area():
    rng = np.random.default_rng(5)
    b = _rand_boxes(rng, 500)
    b[::7, 2] = b[::7, 0]                                          # zero-area boxes
    s = rng.random(500).astype(np.float32)
    keep = _native.nms_tf113(b, s, 500, 0.3)
    assert np.all(np.diff(s[keep]) <= 0)                           # score-sorted
    za = set(range(0, 500, 7))
    assert za.issubset(set(keep.tolist()))                         # zero-area never suppressed
    lib = _native.lib()
    bc = np.ascontiguousarray(b)
    for i in keep[:40]:
        for j in keep[:40]:
            if i != j:
                assert lib.oracle_iou(bc.ctypes.data, int(i), int(j)) <= 0.3


def test_heap_pop_order_ties_is_not_index_order():
    # all-equal scores: libstdc++'s heap does NOT pop in index order — the quirk the CUDA
    # kernels must reproduce (TF 1.13 has no index tie-break).
    order = _native.heap_pop_order(np.ones(16, dtype=np.float32))
    assert sorted(order.tolist()) == list(range(16))
    assert order[0] == 0 and order.tolist() != list(range(16))
    # strictly decreasing scores: index order
    assert _native.heap_pop_order(np.arange(50, 0, -1).astype(np.float32)).tolist() == list(range(50))


def test_crop_and_resize_loop_equals_vectorised():
    rng = np.random.default_rng(3)
    img = rng.normal(size=(2, 9, 11, 5)).astype(np.float32)
    boxes = _rand_boxes(rng, 30)
    boxes[0] = [0, 0, 1, 1]
    boxes[1] = [-0.2, 0.1, 1.3, 0.9]                               # partly outside -> zeros
    boxes[2] = [0.5, 0.5, 0.5, 0.5]                                # zero area
    bi = rng.integers(0, 2, 30)
    a = GL.crop_and_resize(img, boxes, bi, (7, 7))
    b = GL.crop_and_resize_fast(img, boxes, bi, (7, 7))
    assert np.array_equal(a, b)
    # full-image box with crop == image size is the identity
    c = GL.crop_and_resize(img, np.array([[0, 0, 1, 1]], np.float32), [1], (9, 11))
    assert np.allclose(c[0], img[1], atol=1e-6)


def test_crop_and_resize_matches_torch_grid_sample_inside_the_image():
    """Independent cross-check of the tf.image.crop_and_resize restatement (third party, parity unpinned): for boxes inside
    [0, 1] its sampling grid y1*(H-1) + i*(y2-y1)*(H-1)/(P-1) is exactly torch's bilinear grid_sample with
    align_corners=True over the same box, so the two must agree to float32 rounding."""
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(17)
    N, H, W, C, P = 2, 13, 10, 6, 7
    img = rng.normal(size=(N, H, W, C)).astype(np.float32)
    lo = rng.uniform(0.0, 0.6, size=(24, 2))
    boxes = np.concatenate([lo, lo + rng.uniform(0.05, 0.4, size=(24, 2))], axis=1).astype(np.float32)
    boxes = np.clip(boxes, 0, 1)
    boxes[0] = [0, 0, 1, 1]
    bi = rng.integers(0, N, 24)
    got = GL.crop_and_resize_fast(img, boxes, bi, (P, P))
    t = torch.from_numpy(img).permute(0, 3, 1, 2)[torch.from_numpy(bi)]           # [n,C,H,W]
    steps = torch.linspace(0, 1, P, dtype=torch.float64)
    b = torch.from_numpy(boxes).double()
    ys = b[:, 0:1] + steps[None, :] * (b[:, 2:3] - b[:, 0:1])                      # normalised [0,1] sample positions
    xs = b[:, 1:2] + steps[None, :] * (b[:, 3:4] - b[:, 1:2])
    grid = torch.stack([(2 * xs - 1)[:, None, :].expand(-1, P, -1), (2 * ys - 1)[:, :, None].expand(-1, -1, P)], dim=-1)
    want = torch.nn.functional.grid_sample(t.double(), grid, mode="bilinear", padding_mode="border", align_corners=True)
    want = want.permute(0, 2, 3, 1).numpy()
    assert np.abs(got - want).max() < 2e-5


def test_skimage_resize_matches_scipy_map_coordinates_including_borders():
    """Second independent cross-check of the skimage.transform.resize restatement (order 1, mode='constant', cval 0; third
    party, parity unpinned): its coordinate map (out + 0.5) * scale - 0.5 evaluated by scipy.ndimage.map_coordinates(order=1,
    mode="grid-constant", cval=0: the input extended by zeros) — the border rows / columns that blend with cval included, which the cv2 comparison leaves out —
    on the shapes of the detect path: 132 -> 256 upscaling, 28x28 mask -> box, downscaling."""
    ndimage = pytest.importorskip("scipy.ndimage")
    rng = np.random.default_rng(23)
    for (h, w), (oh, ow) in (((132, 132), (256, 256)), ((28, 28), (61, 17)), ((28, 28), (5, 90)), ((40, 64), (20, 32))):
        img = rng.uniform(0.2, 1.0, size=(h, w))
        got = H.skimage_resize(img, (oh, ow))
        r = (np.arange(oh) + 0.5) * (h / oh) - 0.5
        c = (np.arange(ow) + 0.5) * (w / ow) - 0.5
        rr, cc = np.meshgrid(r, c, indexing="ij")
        want = ndimage.map_coordinates(img, [rr, cc], order=1, mode="grid-constant", cval=0.0, prefilter=False)
        want = np.where(want != 0.0, np.clip(want, img.min(), img.max()), want)       # skimage clips to the input range
        assert got.shape == (oh, ow)
        assert np.abs(got - want).max() < 1e-12, ((h, w), (oh, ow), np.abs(got - want).max())


def test_roi_levels_edge_cases():
    boxes = np.array([[[0, 0, 1, 1], [0, 0, 0, 0], [0.1, 0.1, 0.1, 0.5], [0, 0, 224 / 256, 224 / 256],
                       [0, 0, 0.05, 0.05]]], dtype=np.float32)
    lv = GL.roi_levels(boxes, np.float32(256 * 256))
    assert lv.tolist() == [[4, 2, 2, 4, 2]]
    lv = GL.roi_levels(boxes, np.float32(1024 * 1024))
    assert lv.tolist() == [[5, 2, 2, 5, 2]]


def test_skimage_resize_vs_cv2_interior():
    import cv2
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, size=(132, 132), dtype=np.uint8)
    out = H.skimage_resize(img, (256, 256), preserve_range=True)
    ref = cv2.resize(img.astype(np.float64), (256, 256), interpolation=cv2.INTER_LINEAR)
    # identical half-pixel-centre bilinear in the interior; borders differ (cval=0 blend vs replicate)
    assert np.allclose(out[2:-2, 2:-2], ref[2:-2, 2:-2], atol=1e-9)
    assert out[0, 5] < ref[0, 5] + 1e-9
    # 28x28 float mask up-scaling used by unmold_mask
    m = rng.random((28, 28)).astype(np.float32)
    o2 = H.skimage_resize(m, (61, 35))
    r2 = cv2.resize(m.astype(np.float64), (35, 61), interpolation=cv2.INTER_LINEAR)
    assert np.allclose(o2[2:-2, 2:-2], r2[2:-2, 2:-2], atol=1e-9)


def test_zscale_on_shipped_fits(golden_dir):
    for name, nan_count in (("galaxy0002.fits", 288), ("sidelobe0001.fits", 0)):
        raw = open(os.path.join(golden_dir, name), "rb").read()
        data, hdr = H.parse_fits_primary(raw)
        assert data.shape == (132, 132) and hdr["BITPIX"] == -32
        assert int(np.isnan(data).sum()) == nan_count
        rgb = H.fits_to_rgb(data)
        assert rgb.shape == (132, 132, 3) and rgb.dtype == np.uint8
        assert np.array_equal(rgb[..., 0], rgb[..., 1]) and np.array_equal(rgb[..., 1], rgb[..., 2])
        assert rgb.max() == 255 and rgb.min() == 0
        x = np.array(data, dtype=np.float32)
        x[np.isnan(x)] = np.nanmin(x)
        vmin, vmax = H.zscale_limits(x, 0.25)
        assert x.min() <= vmin < vmax <= x.max()


def test_top_k_ties_lower_index_first():
    v = np.array([0.5, 1.0, 0.5, 1.0, 0.25], dtype=np.float32)
    assert GL.top_k_indices(v, 4).tolist() == [1, 3, 0, 2]
