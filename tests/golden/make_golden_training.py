#!/usr/bin/env python
"""Generates tests/golden/training_golden.npz by running the REAL reference training-path host functions.

Runs only in the build container (needs /root/reference; TF / Keras / skimage / astropy are replaced by the inert
stubs of make_golden_from_reference.py, scipy is real).  Executed reference code (pure numpy / scipy bodies only):

  mrcnn/utils.py : compute_overlaps (:147-163), box_refinement (:275-298), resize_mask (:564-583, scipy zoom order 0),
                   extract_bboxes (:49-72), trim_zeros (:715-722)
  mrcnn/model.py : build_rpn_targets (:1536-1644) with np.random seeded — the random sub-sampling is part of the
                   golden, so the restatement must draw from numpy's global generator in the same order;
                   load_image_gt (:1277-1381, no augmentation, scale == 1 and the zoom path) and data_generator
                   (:1721-1904, first batches, shuffle off and on) on an in-memory Dataset subclass;
                   smooth-L1 / loss graphs are TF and stay unpinned.
The outputs pin oracle/train_ops.py and the product's mrcnn/training.py host functions (tests/test_training_host.py).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_from_reference as base  # noqa: E402

OUT = os.path.join(HERE, "training_golden.npz")


def synth_gt(rng, S, n_obj):
    """n_obj elliptical blobs -> masks [S,S,n] bool, class ids [n] int32 in 1..3"""
    yy, xx = np.mgrid[0:S, 0:S]
    masks = np.zeros((S, S, n_obj), dtype=bool)
    for i in range(n_obj):
        cy, cx = rng.uniform(0.1 * S, 0.9 * S, 2)
        ry, rx = rng.uniform(2, 0.12 * S, 2)
        masks[:, :, i] = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0
    return masks, rng.integers(1, 4, n_obj).astype(np.int32)


def main():
    base.install_stubs()
    sys.path.insert(0, base.REF)
    import logging
    logging.disable(logging.CRITICAL)
    if not hasattr(np, "bool"):
        np.bool = bool
    from mrcnn import utils, config as cfgmod, model as modellib

    class C(cfgmod.Config):
        NAME = "golden_train"
        NUM_CLASSES = 4
        GPU_COUNT = 1
        IMAGES_PER_GPU = 2
        IMAGE_MIN_DIM = 128
        IMAGE_MAX_DIM = 128
        RPN_ANCHOR_SCALES = (4, 8, 16, 32, 64)
        MEAN_PIXEL = np.array([0, 0, 0])
        RPN_TRAIN_ANCHORS_PER_IMAGE = 64
        MAX_GT_INSTANCES = 12
        TRAIN_ROIS_PER_IMAGE = 32
        USE_MINI_MASK = False
    cfg = C()
    g = {}
    rng = np.random.default_rng(11)

    # ---- compute_overlaps / box_refinement ------------------------------------------------------
    b1 = np.sort(rng.integers(0, 128, (40, 4)), axis=1).astype(np.float64)[:, [0, 1, 2, 3]]
    b1 = np.stack([b1[:, 0], b1[:, 1], b1[:, 2] + 1, b1[:, 3] + 1], 1)
    b2 = np.stack([b1[:9, 0] + 2, b1[:9, 1] - 1, b1[:9, 2] + 5, b1[:9, 3] + 3], 1)
    g["ov_b1"], g["ov_b2"] = b1, b2
    g["ov_out"] = utils.compute_overlaps(b1, b2)
    g["refine_out"] = utils.box_refinement(b1[:9], b2)

    # ---- build_rpn_targets ------------------------------------------------------------------------
    shapes = modellib.compute_backbone_shapes(cfg, cfg.IMAGE_SHAPE)
    anchors = utils.generate_pyramid_anchors(cfg.RPN_ANCHOR_SCALES, cfg.RPN_ANCHOR_RATIOS, shapes,
                                             cfg.BACKBONE_STRIDES, cfg.RPN_ANCHOR_STRIDE)
    g["rpn_anchors"] = anchors
    for tag, n_obj, crowd in (("a", 5, False), ("b", 11, True), ("c", 1, False)):
        masks, cls = synth_gt(rng, 128, n_obj)
        keep = masks.sum(axis=(0, 1)) > 0
        masks, cls = masks[:, :, keep], cls[keep]
        boxes = utils.extract_bboxes(masks)
        if crowd:
            cls = cls.copy()
            cls[::4] *= -1
        np.random.seed(100 + n_obj)
        match, bbox = modellib.build_rpn_targets((128, 128, 3), anchors, cls, boxes, cfg)
        g["rpn_%s_cls" % tag], g["rpn_%s_boxes" % tag] = cls, boxes
        g["rpn_%s_seed" % tag] = np.array([100 + n_obj])
        g["rpn_%s_match" % tag], g["rpn_%s_bbox" % tag] = match, bbox
        g["rpn_%s_masks" % tag] = np.packbits(masks)
        g["rpn_%s_masks_shape" % tag] = np.array(masks.shape)

    # ---- resize_mask (scipy zoom, order 0) ------------------------------------------------------------
    m, _ = synth_gt(rng, 100, 3)
    m = m[:, :80]
    g["rm_in"] = np.packbits(m)
    g["rm_in_shape"] = np.array(m.shape)
    out = utils.resize_mask(m, 1.28, [(0, 0), (13, 13), (0, 0)])
    g["rm_out"] = np.packbits(out)
    g["rm_out_shape"] = np.array(out.shape)

    # ---- Dataset + load_image_gt + data_generator -------------------------------------------------------
    class DS(utils.Dataset):
        def __init__(self, n):
            super().__init__()
            for i, name in enumerate(["sidelobe", "source", "galaxy"]):
                self.add_class("rg", i + 1, name)
            r = np.random.default_rng(5)
            self.items = []
            for i in range(n):
                img = r.integers(0, 256, (128, 128, 3), dtype=np.uint8)
                masks, cls = synth_gt(r, 128, int(r.integers(1, 8)))
                self.items.append((img, masks, cls))
                self.add_image("rg", image_id=i, path="mem://%d" % i)

        def load_image(self, image_id):
            return self.items[image_id][0]

        def load_mask(self, image_id):
            return self.items[image_id][1], self.items[image_id][2]

    ds = DS(6)
    ds.prepare()
    g["ds_class_ids"] = np.array(ds.class_ids)
    g["ds_source_class_ids_rg"] = np.array(ds.source_class_ids["rg"])
    for i in range(6):
        g["ds_img_%d" % i] = ds.items[i][0]
        g["ds_masks_%d" % i] = np.packbits(ds.items[i][1])
        g["ds_masks_shape_%d" % i] = np.array(ds.items[i][1].shape)
        g["ds_cls_%d" % i] = ds.items[i][2]
    image, meta, cls, bbox, mask = modellib.load_image_gt(ds, cfg, 2, use_mini_mask=False)
    g["gt2_image"], g["gt2_meta"], g["gt2_cls"], g["gt2_bbox"] = image, meta, cls, bbox
    g["gt2_mask"] = np.packbits(mask)
    g["gt2_mask_shape"] = np.array(mask.shape)
    np.random.seed(77)
    gen = modellib.data_generator(ds, cfg, shuffle=False, batch_size=cfg.BATCH_SIZE)
    for step in range(2):
        inputs, outputs = next(gen)
        assert outputs == []
        for k, arr in enumerate(inputs):
            a = np.asarray(arr)
            g["gen_s%d_in%d" % (step, k)] = np.packbits(a) if a.dtype == bool else a
            g["gen_s%d_in%d_shape" % (step, k)] = np.array(a.shape)
    np.random.seed(78)
    gen = modellib.data_generator(ds, cfg, shuffle=True, batch_size=cfg.BATCH_SIZE)
    inputs, _ = next(gen)
    g["gen_shuffle_meta"] = inputs[1]
    g["gen_shuffle_match_sum"] = np.array([int((inputs[2] == 1).sum()), int((inputs[2] == -1).sum())])
    g["gen_shuffle_rpn_bbox"] = inputs[3]

    np.savez_compressed(OUT, **g)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(g), "arrays")


if __name__ == "__main__":
    main()
