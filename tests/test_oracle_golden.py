"""CPU: the oracle's pure-numpy host pieces against vectors produced by the REAL reference
functions (tests/golden/make_golden_from_reference.py ran mrcnn/utils.py + mrcnn/model.py from
/root/reference with TF/Keras/astropy/skimage stubbed)."""
import numpy as np

from oracle import host_ops as H


def test_anchors_match_reference(golden):
    for S, scales in ((256, (4, 8, 16, 32, 64)), (128, (4, 8, 16, 32, 64))):
        shapes = H.compute_backbone_shapes((S, S, 3))
        assert np.array_equal(shapes, golden["backbone_shapes_%d" % S])
        a = H.generate_pyramid_anchors(scales, [0.5, 1, 2], shapes, [4, 8, 16, 32, 64], 1)
        assert np.array_equal(a, golden["anchors_px_%d" % S])
        n = H.norm_boxes(a, (S, S))
        assert n.dtype == np.float32
        assert np.array_equal(n, golden["anchors_norm_%d" % S])
    a = H.generate_pyramid_anchors((32, 64, 128, 256, 512), [0.5, 1, 2],
                                   H.compute_backbone_shapes((1024, 1024, 3)), [4, 8, 16, 32, 64], 1)
    assert a.shape[0] == int(golden["anchors_1024_count"][0]) == 261888
    assert np.array_equal(a[::97], golden["anchors_px_1024_sub"])
    assert np.array_equal(H.norm_boxes(a, (1024, 1024))[::97], golden["anchors_norm_1024_sub"])
    assert np.array_equal(a.sum(axis=0), golden["anchors_px_1024_colsum"])
    assert np.array_equal(H.generate_anchors(32, [0.5, 1, 2], [3, 5], 16, 2), golden["gen_anchors_small"])
    assert np.array_equal(H.get_anchors((256, 256, 3), (4, 8, 16, 32, 64)), golden["get_anchors_256"])


def test_norm_denorm_match_reference(golden):
    assert np.array_equal(H.norm_boxes(golden["norm_in"], (132, 200)), golden["norm_out"])
    out = H.denorm_boxes(golden["denorm_in"], (132, 132))
    assert out.dtype == np.int32
    assert np.array_equal(out, golden["denorm_out"])


def test_resize_image_bookkeeping_and_mold(golden):
    img = golden["resize_in"]
    out, window, scale, padding, crop = H.resize_image(img, min_dim=128, max_dim=128, min_scale=0, mode="square")
    assert out.dtype == np.uint8 and np.array_equal(out, golden["resize_out"])
    assert tuple(window) == tuple(golden["resize_window"])
    assert float(scale) == float(golden["resize_scale"][0])
    assert np.array_equal(np.array(padding), golden["resize_padding"])
    molded, metas, windows = H.mold_inputs([img], min_dim=128, max_dim=128, min_scale=0, mode="square",
                                           mean_pixel=np.array([0, 0, 0]), num_classes=4)
    assert np.array_equal(molded, golden["mold_molded"].astype(np.float32))
    assert np.array_equal(metas, golden["mold_metas"])
    assert np.array_equal(windows, golden["mold_windows"])
    meta = H.compose_image_meta(3, (132, 132, 3), (256, 256, 3), (0, 0, 256, 256), 1.9393939,
                                np.zeros([4], dtype=np.int32))
    assert np.array_equal(meta, golden["compose_meta"])


def test_gray2rgb_rounding(golden):
    chans = golden["gray_in"]
    u8 = np.stack([np.array((c * np.float32(255)).round(), dtype=np.uint8) for c in chans], axis=-1)
    assert np.array_equal(u8, golden["gray2rgb_u8"])
    assert np.array_equal((chans[1] / np.max(chans[1])).astype(np.float32), golden["norm_img_out"])


def test_unmold_detections_logic(golden, monkeypatch):
    """Box arithmetic, zero-area filtering and paste are pinned by the reference run in which
    skimage.resize was replaced by a nearest-neighbour stub; use the same stub here."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location(
        "mkgold", os.path.join(os.path.dirname(__file__), "golden", "make_golden_from_reference.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    monkeypatch.setattr(H, "skimage_resize", lambda m, shape, **kw: mk.nearest_resize(m, shape))
    masks = golden["unmold_mrcnn_mask"].astype(np.float32)
    for tag in "ab":
        args = golden["unmold_%s_args" % tag]
        orig, molded_shape, window = tuple(args[:3]), tuple(args[3:6]), np.array(args[6:10])
        b, ci, sc, fm = H.unmold_detections(golden["unmold_%s_det" % tag], masks, orig, molded_shape, window)
        assert b.dtype == np.int32 and np.array_equal(b, golden["unmold_%s_boxes" % tag])
        assert np.array_equal(ci, golden["unmold_%s_class_ids" % tag])
        assert np.array_equal(sc, golden["unmold_%s_scores" % tag])
        assert tuple(fm.shape) == tuple(golden["unmold_%s_masks_shape" % tag])
        assert np.array_equal(np.packbits(fm.astype(np.uint8)), golden["unmold_%s_masks" % tag])
