"""Host pieces of `run.py train` (reference scripts/run.py:246-989): SourceDataset loaders, mask reading, the
train / cross-validation split and argument validation.  No GPU (images are only read by the GPU tests)."""
import importlib.util
import json
import os

import numpy as np


def _run_module():
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "caesar-mrcnn_b200", "scripts", "run.py")
    spec = importlib.util.spec_from_file_location("run_b200_train", path)
    run = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(run)
    return run


def _write_dataset(tmp_path, n_img=12):
    from mrcnn import fitsio
    rng = np.random.default_rng(0)
    lines, jsons = [], []
    for i in range(n_img):
        img = rng.normal(0, 1e-3, (64, 64)).astype(np.float32)
        ipath = str(tmp_path / ("img%d.fits" % i))
        fitsio.write_primary(ipath, img)
        objs = []
        for k in range(2):
            m = np.zeros((64, 64), np.float32)
            m[5 + 20 * k:15 + 20 * k, 8 + i:20 + i] = 1
            if k == 1:
                m[0, 0] = np.nan                       # NaN -> minimum (0) in the raw read path
            mpath = str(tmp_path / ("mask%d_%d.fits" % (i, k)))
            fitsio.write_primary(mpath, m)
            lines.append("%s,%s,%s" % (ipath, mpath, "source" if k == 0 else "galaxy"))
            objs.append({"mask": os.path.basename(mpath), "class": "source" if k == 0 else "galaxy", "nislands": 1,
                         "sidelobe-mixed": 0, "sidelobe-near": 0})
        jpath = str(tmp_path / ("img%d.json" % i))
        with open(jpath, "w") as f:
            json.dump({"img": os.path.basename(ipath), "objs": objs, "telescope": "t", "bkg": 0, "rms": 1, "bmaj": 1, "bmin": 1,
                       "dx": 1, "dy": 1, "nx": 64, "ny": 64}, f)
        jsons.append(jpath)
    lst = str(tmp_path / "list.dat")
    open(lst, "w").write("\n".join(lines) + "\n")
    jl = str(tmp_path / "jlist.dat")
    open(jl, "w").write("\n".join(jsons) + "\n")
    return lst, jl


def test_source_dataset_loaders_and_masks(tmp_path):
    run = _run_module()
    lst, jl = _write_dataset(tmp_path)
    ds = run.SourceDataset()
    assert ds.set_class_dict('{"sidelobe":1,"source":2,"galaxy":3}') == 0 and ds.nclasses == 4
    assert ds.load_data_from_list(lst, nmaximgs=-1) == 0 and ds.loaded_imgs == 24
    ds.prepare()
    assert ds.num_classes == 4 and ds.class_names == ["BG", "sidelobe", "source", "galaxy"]
    mask, cls = ds.load_mask(1)                                   # second line: the galaxy object of image 0
    assert mask.shape == (64, 64, 1) and mask.dtype == bool and cls.tolist() == [3] and cls.dtype == np.int32
    assert mask[:, :, 0].sum() == 10 * 12 and not mask[0, 0, 0]
    js = run.SourceDataset()
    js.set_class_dict('{"sidelobe":1,"source":2,"galaxy":3}')
    assert js.load_data_from_json_list(jl, -1) == 0 and js.loaded_imgs == 12
    js.prepare()
    mask, cls = js.load_mask(3)
    assert mask.shape == (64, 64, 2) and cls.tolist() == [2, 3]
    assert js.nobjs_per_class[2] == 12 and js.nobjs_per_class[3] == 12
    assert js.image_info[3]["sidelobes_mixed_or_near"] == [0, 0] and js.image_reference(3).endswith("img3.fits")
    bad = run.SourceDataset()
    bad.set_class_dict('{"source":2}')
    assert bad.load_data_from_list(lst) == 0 and bad.loaded_imgs == 12       # the galaxy lines are skipped
    assert run.SourceDataset().set_class_dict("") == -1


def test_train_val_split_and_validation(tmp_path):
    import random
    run = _run_module()
    lst, jl = _write_dataset(tmp_path)
    random.seed(3)
    tr, va = run.create_train_val_sets_from_filelist(lst, 0.25, str(tmp_path / "tr.dat"), str(tmp_path / "va.dat"))
    a, b = open(tr).read().split(), open(va).read().split()
    assert len(a) == 18 and len(b) == 6 and sorted(a + b) == sorted(open(lst).read().split())
    assert run.main(["train"]) == 1                                               # no datalist
    assert run.main(["train", "--datalist", lst, "--weight_classes"]) == 1
    assert run.main(["train", "--datalist", lst, "--backbone", "resnet50"]) == 1
    args = run.parse_args(["train", "--datalist", lst, "--nimg_per_gpu", "2", "--no_rpn_bbox_loss", "--mrcnn_mask_loss_weight", "2.5"])
    assert run.validate_args(args) == 0
    cfg = run.make_config(args)
    assert cfg.BATCH_SIZE == 2 and cfg.TRAIN_ROIS_PER_IMAGE == 512 and cfg.RPN_TRAIN_ANCHORS_PER_IMAGE == 512
    assert cfg.MAX_GT_INSTANCES == 300 and cfg.USE_MINI_MASK is False and cfg.LEARNING_RATE == 0.0005
    assert cfg.USE_LOSSES["rpn_bbox_loss"] is False and cfg.USE_LOSSES["rpn_class_loss"] is True
    assert cfg.LOSS_WEIGHTS["mrcnn_mask_loss"] == 2.5
