#!/usr/bin/env python
"""Generates tests/golden/ref_numpy_golden.npz by running the REAL reference functions.

Runs only in the build container (needs /root/reference).  TensorFlow, Keras, astropy,
scikit-image, matplotlib and distutils are absent, so they are replaced by inert stub modules
before `mrcnn.utils` / `mrcnn.model` / `mrcnn.config` are imported from /root/reference; only
functions whose bodies are pure numpy are then executed:

  mrcnn/utils.py : generate_anchors, generate_pyramid_anchors, norm_boxes, denorm_boxes,
                   resize_image (scale == 1 path: padding/window bookkeeping), normalize_img,
                   gray2rgb, compute_iou, extract_bboxes
  mrcnn/model.py : compute_backbone_shapes, compose_image_meta, parse_image_meta, mold_image,
                   MaskRCNN.get_anchors, MaskRCNN.mold_inputs (scale == 1),
                   MaskRCNN.unmold_detections (with skimage.transform.resize stubbed by a
                   nearest-neighbour resampler supplied here and recorded in the file, so the box
                   arithmetic / filtering / paste logic is pinned independently of the resize)
  mrcnn/config.py: Config defaults and derived attributes

The outputs are the golden vectors the oracle (oracle/host_ops.py) is pinned against in
tests/test_oracle_golden.py.  Nothing here is imported by the product.
"""
import os
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_numpy_golden.npz")


class _Anything:
    """Class usable as a base class, callable, and attribute sink."""
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything


def _stub(name, **attrs):
    m = _StubModule(name)
    m.__path__ = []
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def nearest_resize(image, output_shape, **kw):
    """Stand-in for skimage.transform.resize used ONLY while generating unmold goldens."""
    image = np.asarray(image, dtype=np.float64)
    rows, cols = int(output_shape[0]), int(output_shape[1])
    r = np.minimum((np.arange(rows) + 0.5) * image.shape[0] / max(rows, 1), image.shape[0] - 1).astype(int)
    c = np.minimum((np.arange(cols) + 0.5) * image.shape[1] / max(cols, 1), image.shape[1] - 1).astype(int)
    return image[r][:, c] if rows and cols else np.zeros((rows, cols))


def install_stubs():
    class LooseVersion:
        def __init__(self, v):
            self.v = tuple(int(p) for p in str(v).split(".")[:3] if p.isdigit())

        def __ge__(self, o):
            return self.v >= o.v

        def __lt__(self, o):
            return self.v < o.v

    _stub("distutils")
    _stub("distutils.version", LooseVersion=LooseVersion)
    tf = _stub("tensorflow", __version__="1.13.2")
    _stub("tensorflow.python")
    k = _stub("keras", __version__="2.2.4")
    for sub in ("backend", "layers", "engine", "models", "callbacks", "optimizers", "regularizers",
                "utils", "utils.data_utils", "engine.saving", "engine.topology"):
        setattr(k, sub.split(".")[0], _stub("keras." + sub))
    sk = _stub("skimage", __version__="0.15.0")
    sk.transform = _stub("skimage.transform", resize=nearest_resize)
    sk.color = _stub("skimage.color")
    sk.io = _stub("skimage.io")
    sk.measure = _stub("skimage.measure")
    for name in ("astropy", "astropy.io", "astropy.io.ascii", "astropy.io.fits", "astropy.units",
                 "astropy.modeling", "astropy.modeling.parameters", "astropy.modeling.core",
                 "astropy.wcs", "astropy.visualization", "astropy.stats", "astropy.coordinates",
                 "matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.lines",
                 "imgaug", "imgaug.augmenters", "h5py", "IPython", "IPython.display"):
        _stub(name)
    return tf


def main():
    install_stubs()
    sys.path.insert(0, REF)
    import logging
    logging.disable(logging.CRITICAL)
    from mrcnn import utils, config as cfgmod, model as modellib
    # numpy >= 1.24 dropped np.bool, which utils.unmold_mask still spells
    if not hasattr(np, "bool"):
        np.bool = bool

    g = {}
    # ---- Config -----------------------------------------------------------------------------
    class C(cfgmod.Config):
        NAME = "golden"
        NUM_CLASSES = 4
        GPU_COUNT = 1
        IMAGES_PER_GPU = 1
        IMAGE_MIN_DIM = 256
        IMAGE_MAX_DIM = 256
        RPN_ANCHOR_SCALES = (4, 8, 16, 32, 64)
        MEAN_PIXEL = np.array([0, 0, 0])
        DETECTION_MIN_CONFIDENCE = 0
    c = C()
    base = cfgmod.Config()
    names = sorted(a for a in dir(base) if not a.startswith("__") and not callable(getattr(base, a)))
    g["config_attr_names"] = np.array(names)
    g["config_attr_reprs"] = np.array([repr(getattr(base, a)) for a in names])
    g["config_derived"] = np.array([c.BATCH_SIZE, c.IMAGE_META_SIZE] + list(c.IMAGE_SHAPE))

    # ---- anchors ----------------------------------------------------------------------------
    for S, scales in ((256, (4, 8, 16, 32, 64)), (1024, (32, 64, 128, 256, 512)), (128, (4, 8, 16, 32, 64))):
        shapes = modellib.compute_backbone_shapes(c, (S, S, 3))
        a = utils.generate_pyramid_anchors(scales, [0.5, 1, 2], shapes, [4, 8, 16, 32, 64], 1)
        if S == 1024:       # 261 888 anchors: keep every 97th row + column sums (file size)
            g["anchors_px_1024_sub"] = a[::97]
            g["anchors_norm_1024_sub"] = utils.norm_boxes(a, (S, S))[::97]
            g["anchors_px_1024_colsum"] = a.sum(axis=0)
            g["anchors_1024_count"] = np.array([a.shape[0]])
        else:
            g["anchors_px_%d" % S] = a.astype(np.float32) if np.all(a == a.astype(np.float32)) else a
            g["anchors_norm_%d" % S] = utils.norm_boxes(a, (S, S))
        g["backbone_shapes_%d" % S] = shapes
    g["gen_anchors_small"] = utils.generate_anchors(32, [0.5, 1, 2], [3, 5], 16, 2)

    # ---- MaskRCNN.get_anchors through the real method (fake self) ----------------------------
    fake = types.SimpleNamespace(config=c)
    g["get_anchors_256"] = modellib.MaskRCNN.get_anchors(fake, (256, 256, 3))

    # ---- norm / denorm ----------------------------------------------------------------------
    rng = np.random.default_rng(7)
    boxes_px = rng.integers(0, 200, size=(64, 4)).astype(np.float64)
    g["norm_in"] = boxes_px
    g["norm_out"] = utils.norm_boxes(boxes_px, (132, 200))
    bn = rng.random((257, 4)).astype(np.float32)
    bn[:16] = (np.arange(16)[:, None] + 0.5) / 131.0      # exact .5 cases for half-to-even
    g["denorm_in"] = bn
    g["denorm_out"] = utils.denorm_boxes(bn, (132, 132))

    # ---- resize_image bookkeeping (scale == 1: no skimage) + mold_inputs ----------------------
    img = rng.integers(0, 256, size=(128, 100, 3), dtype=np.uint8)
    out, window, scale, padding, crop = utils.resize_image(img, min_dim=128, max_dim=128, min_scale=0, mode="square")
    g["resize_in"] = img
    g["resize_out"] = out
    g["resize_window"] = np.array(window)
    g["resize_scale"] = np.array([scale], dtype=np.float64)
    g["resize_padding"] = np.array(padding)
    c128 = C()
    c128.IMAGE_MIN_DIM = c128.IMAGE_MAX_DIM = 128
    molded, metas, windows = modellib.MaskRCNN.mold_inputs(types.SimpleNamespace(config=c128), [img])
    g["mold_molded"] = molded
    g["mold_metas"] = metas
    g["mold_windows"] = windows
    g["compose_meta"] = modellib.compose_image_meta(3, (132, 132, 3), (256, 256, 3), (0, 0, 256, 256), 1.9393939,
                                                   np.zeros([4], dtype=np.int32))
    g["mold_image_f"] = modellib.mold_image(img[:4, :4], C())

    # ---- normalize_img / gray2rgb -------------------------------------------------------------
    chans = [np.clip(rng.normal(0.3, 0.3, (33, 17)), 0, 1).astype(np.float32) for _ in range(3)]
    chans[0][0, :5] = np.array([0.5 / 255, 1.5 / 255, 2.5 / 255, 3.5 / 255, 254.5 / 255], dtype=np.float32)
    g["gray_in"] = np.stack(chans)
    g["norm_img_out"] = utils.normalize_img(chans[1])
    g["gray2rgb_u8"] = utils.gray2rgb(chans, True)
    g["gray2rgb_f32"] = utils.gray2rgb(chans, False)

    # ---- unmold_detections with the nearest-neighbour resize stub -----------------------------
    n = 37
    masks = rng.random((40, 28, 28, 4)).astype(np.float16)
    g["unmold_mrcnn_mask"] = masks
    for tag, orig, molded_shape, window in (("a", (132, 132, 3), (256, 256, 3), (0, 0, 256, 256)),
                                            ("b", (100, 180, 3), (256, 256, 3), (57, 0, 199, 256))):
        wn = utils.norm_boxes(np.array(window), molded_shape[:2])       # detections live inside it
        det = np.zeros((40, 6), dtype=np.float32)
        fy = np.sort(rng.random((n, 2)).astype(np.float32), axis=1)
        fx = np.sort(rng.random((n, 2)).astype(np.float32), axis=1)
        det[:n, 0] = wn[0] + fy[:, 0] * (wn[2] - wn[0])
        det[:n, 2] = wn[0] + fy[:, 1] * (wn[2] - wn[0])
        det[:n, 1] = wn[1] + fx[:, 0] * (wn[3] - wn[1])
        det[:n, 3] = wn[1] + fx[:, 1] * (wn[3] - wn[1])
        det[:n, :4] = np.clip(det[:n, :4], [wn[0], wn[1], wn[0], wn[1]], [wn[2], wn[3], wn[2], wn[3]])
        det[5, 2] = det[5, 0] - np.float32(0.02)            # negative height -> filtered on host (model.py:2603-2610)
        det[:n, 4] = rng.integers(1, 4, n)
        det[:n, 5] = np.sort(rng.random(n).astype(np.float32))[::-1]
        b, ci, sc, fm = modellib.MaskRCNN.unmold_detections(fake, det, masks.astype(np.float32), orig,
                                                            molded_shape, np.array(window))
        g["unmold_%s_det" % tag] = det
        g["unmold_%s_boxes" % tag] = b
        g["unmold_%s_class_ids" % tag] = ci
        g["unmold_%s_scores" % tag] = sc
        g["unmold_%s_masks" % tag] = np.packbits(fm.astype(np.uint8))
        g["unmold_%s_masks_shape" % tag] = np.array(fm.shape)
        g["unmold_%s_args" % tag] = np.array(list(orig) + list(molded_shape) + list(window))

    # ---- compute_iou / extract_bboxes ---------------------------------------------------------
    bx = np.sort(rng.integers(0, 100, size=(20, 4)), axis=1)[:, [0, 1, 2, 3]].astype(np.float64)
    bx = np.stack([bx[:, 0], bx[:, 1], bx[:, 2] + 1, bx[:, 3] + 1], axis=1)
    area = (bx[:, 2] - bx[:, 0]) * (bx[:, 3] - bx[:, 1])
    g["iou_boxes"] = bx
    g["iou_row0"] = utils.compute_iou(bx[0], bx, area[0], area)

    np.savez_compressed(OUT, **g)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(g), "arrays")


if __name__ == "__main__":
    main()
