"""Host half of the training path (data_generator / load_image_gt / build_rpn_targets and their utils helpers) against
outputs of the REFERENCE's own functions (tests/golden/make_golden_training.py -> training_golden.npz).  No GPU."""
import os

import numpy as np
import pytest

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "training_golden.npz"))


def _cfg():
    from mrcnn.config import Config

    class C(Config):
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
    return C()


def _unpack(name):
    shape = tuple(GOLD[name + "_shape"])
    return np.unpackbits(GOLD[name])[:int(np.prod(shape))].reshape(shape).astype(bool)


def _impls():
    from mrcnn import model as product_model, utils as product_utils
    from oracle import train_ops as oracle_ops
    return [("product", product_utils.compute_overlaps, product_utils.box_refinement, product_model.build_rpn_targets),
            ("oracle", oracle_ops.compute_overlaps, oracle_ops.box_refinement, oracle_ops.build_rpn_targets)]


@pytest.mark.parametrize("which", [0, 1])
def test_overlaps_and_box_refinement_equal_reference(which):
    _, overlaps, refinement, _ = _impls()[which]
    assert np.array_equal(overlaps(GOLD["ov_b1"], GOLD["ov_b2"]), GOLD["ov_out"])
    out = refinement(GOLD["ov_b1"][:9], GOLD["ov_b2"])
    assert out.dtype == GOLD["refine_out"].dtype and np.array_equal(out, GOLD["refine_out"])


@pytest.mark.parametrize("which", [0, 1])
@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_build_rpn_targets_equal_reference_with_the_same_seed(which, tag):
    _, _, _, build = _impls()[which]
    np.random.seed(int(GOLD["rpn_%s_seed" % tag][0]))
    match, bbox = build((128, 128, 3), GOLD["rpn_anchors"], GOLD["rpn_%s_cls" % tag], GOLD["rpn_%s_boxes" % tag], _cfg())
    assert match.dtype == np.int32 and np.array_equal(match, GOLD["rpn_%s_match" % tag])
    assert bbox.dtype == np.float64 and np.array_equal(bbox, GOLD["rpn_%s_bbox" % tag])
    assert (match == 1).sum() <= 32 and (match != 0).sum() <= 64


def test_resize_mask_equal_reference():
    from mrcnn import utils
    out = utils.resize_mask(_unpack("rm_in"), 1.28, [(0, 0), (13, 13), (0, 0)])
    assert np.array_equal(out, _unpack("rm_out"))


class _DS(object):
    """the in-memory dataset of the golden generator, rebuilt on the product's Dataset base class"""
    def __new__(cls):
        from mrcnn import utils

        class DS(utils.Dataset):
            def __init__(self):
                super().__init__()
                for i, name in enumerate(["sidelobe", "source", "galaxy"]):
                    self.add_class("rg", i + 1, name)
                for i in range(6):
                    self.add_image("rg", image_id=i, path="mem://%d" % i)

            def load_image(self, image_id):
                return GOLD["ds_img_%d" % image_id]

            def load_mask(self, image_id):
                shape = tuple(GOLD["ds_masks_shape_%d" % image_id])
                bits = np.unpackbits(GOLD["ds_masks_%d" % image_id])[:int(np.prod(shape))]
                return bits.reshape(shape).astype(bool), GOLD["ds_cls_%d" % image_id]
        ds = DS()
        ds.prepare()
        return ds


def test_dataset_bookkeeping_and_load_image_gt_equal_reference():
    from mrcnn import model as modellib
    ds = _DS()
    assert np.array_equal(ds.class_ids, GOLD["ds_class_ids"])
    assert np.array_equal(np.array(ds.source_class_ids["rg"]), GOLD["ds_source_class_ids_rg"])
    image, meta, cls, bbox, mask = modellib.load_image_gt(ds, _cfg(), 2, use_mini_mask=False)
    assert np.array_equal(image, GOLD["gt2_image"]) and np.array_equal(meta, GOLD["gt2_meta"])
    assert np.array_equal(cls, GOLD["gt2_cls"]) and np.array_equal(bbox, GOLD["gt2_bbox"]) and bbox.dtype == GOLD["gt2_bbox"].dtype
    assert np.array_equal(mask, _unpack("gt2_mask"))


def test_data_generator_batches_equal_reference():
    from mrcnn import model as modellib
    cfg = _cfg()
    np.random.seed(77)
    gen = modellib.data_generator(_DS(), cfg, shuffle=False, batch_size=cfg.BATCH_SIZE)
    for step in range(2):
        inputs, outputs = next(gen)
        assert outputs == [] and len(inputs) == 7
        for k, arr in enumerate(inputs):
            want = GOLD["gen_s%d_in%d" % (step, k)]
            shape = tuple(GOLD["gen_s%d_in%d_shape" % (step, k)])
            if arr.dtype == bool:
                want = np.unpackbits(want)[:int(np.prod(shape))].reshape(shape).astype(bool)
            assert arr.shape == shape and arr.dtype == want.dtype, (step, k, arr.dtype, want.dtype)
            assert np.array_equal(arr, want), (step, k)
    np.random.seed(78)
    inputs, _ = next(modellib.data_generator(_DS(), cfg, shuffle=True, batch_size=cfg.BATCH_SIZE))
    assert np.array_equal(inputs[1], GOLD["gen_shuffle_meta"]) and np.array_equal(inputs[3], GOLD["gen_shuffle_rpn_bbox"])
    assert [int((inputs[2] == 1).sum()), int((inputs[2] == -1).sum())] == GOLD["gen_shuffle_match_sum"].tolist()
