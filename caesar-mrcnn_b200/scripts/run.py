#!/usr/bin/env python
"""run.py — detect/test/train CLI of the B200 build, flag-compatible with the corresponding subset of the
reference's scripts/run.py (parse_args :1263-1384, InferenceConfig :1652-1706, detect :1172-1189, train :1052-1125,
SourceDataset :246-818).

  run.py detect --image data/galaxy0002.fits --weights share/mrcnn_weights.h5 [--imgsize 256 ...]
  run.py test   --datalist images.txt        --weights share/mrcnn_weights.h5
  run.py train  --datalist train.txt [--dataloader datalist|datalist_json] --nepochs 10 --nimg_per_gpu 2 [--weights w.h5]
                (one process per GPU under torchrun: every rank trains on its own shuffled stream of the dataset and the
                 gradients are averaged over NVLink — the replacement of --ngpu N / ParallelModel)

`detect` reads the FITS image (zscale + uint8 RGB on the GPU) and hands it to `Analyzer.predict` exactly as the
reference's SFinder.run does (mrcnn/sfinder.py:485-493): MaskRCNN.detect, then extract_det_masks (score filter,
merging of connected same-class masks, best-of-overlapping selection) on the GPU, then the reference's JSON catalogue
(`out_<image>.json` or --detect_outfile_json; keys name,x1,x2,y1,y2,class_id,class_name,score,pixels,vertexes,edge).
`test` writes the raw detections with score >= --scoreThr of every FITS file of a list (batched by --nimg_per_gpu).
With --split_img_in_tiles the image is cut into --tile_xsize x --tile_ysize tiles (SFinder.run_parallel,
mrcnn/sfinder.py:549-640): the tiles go through the detector in batches of --nimg_per_gpu, sources cut by tile borders
are merged, and `catalog_<image>.json` is written; launched under torchrun, every rank (GPU) takes the tiles the
reference's MPI ranks would. `train` builds the reference's SourceDataset (image,mask,label lists or per-image JSON files),
splits it into training / cross-validation sets and calls MaskRCNN.train(layers='all') (mrcnn/training.py).  PNG / DS9
output, the ground-truth metrics of `test`, source parameters (WCS / flux), imgaug augmentation and class weights are
outside the path rebuilt here and are rejected or skipped.
Returns exit code 0 on success, 1 on failure, like the reference's main().
"""
import argparse
import json
import logging
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from mrcnn import logger  # noqa: E402
from mrcnn import model as modellib, utils  # noqa: E402
from mrcnn.analyze import Analyzer  # noqa: E402
from mrcnn.sfinder import SFinder  # noqa: E402
from mrcnn.config import Config  # noqa: E402


class SDetectorConfig(Config):
    """Effective `run.py detect/test` configuration of the reference (scripts/run.py:93-239)."""
    NAME = "rg-dataset"
    GPU_COUNT = 1
    IMAGES_PER_GPU = 1
    NUM_CLASSES = 1
    CLASS_NAMES = ["bkg"]
    DETECTION_MIN_CONFIDENCE = 0
    DETECTION_NMS_THRESHOLD = 0.3
    RPN_ANCHOR_SCALES = (4, 8, 16, 32, 64)
    MAX_GT_INSTANCES = 300
    BACKBONE = "resnet101"
    IMAGE_RESIZE_MODE = "square"
    IMAGE_MIN_DIM = 256
    IMAGE_MAX_DIM = 256
    MEAN_PIXEL = np.array([0, 0, 0])
    RPN_NMS_THRESHOLD = 0.7
    ZSCALE_STRETCH = True
    ZSCALE_CONTRASTS = [0.25, 0.25, 0.25]
    NORMALIZE_IMG = True
    IMG_TO_UINT8 = True
    IMG_TO_RGB = True
    BIAS_CONTRAST_STRETCH = False
    IOU_THR = 0.6
    SCORE_THR = 0.7


class SourceDataset(utils.Dataset):
    """reference: scripts/run.py:246-818 — radio-source dataset: one FITS image + one FITS mask file per object.
    Loaders: `image.fits,mask.fits,class_name` lines (load_data_from_list) or per-image JSON files listing the object
    masks (load_data_from_json_file / _list).  Images go through read_fits (zscale + uint8 RGB on the GPU), masks are read
    raw and cast to bool."""

    def __init__(self):
        utils.Dataset.__init__(self)
        self.class_id_map = {}
        self.nclasses = 0
        self.loaded_imgs = 0
        self.zscale_contrasts = [0.25, 0.25, 0.25]
        self.nobjs_per_class = {}

    def set_class_dict(self, class_dict_str):
        if class_dict_str == "":
            logger.error("Empty string given!")
            return -1
        try:
            class_dict = json.loads(class_dict_str)
        except Exception:
            logger.error("Failed to get dictionary from string!")
            return -1
        self.class_id_map = class_dict
        for class_name in self.class_id_map:
            class_id = self.class_id_map[class_name]
            self.add_class("rg-dataset", class_id, class_name)
            self.nobjs_per_class[class_id] = 0
        self.class_id_map['bkg'] = 0
        self.nobjs_per_class[0] = 0
        self.nclasses = len(self.class_id_map)
        return 0

    def _add(self, image_path, mask_paths, class_ids, **extra):
        import uuid
        self.add_image("rg-dataset", image_id=str(uuid.uuid1()), path=image_path, path_masks=mask_paths, class_ids=class_ids, **extra)
        for cid in class_ids:
            self.nobjs_per_class[cid] = self.nobjs_per_class.get(cid, 0) + 1
        self.loaded_imgs += 1

    def load_data_from_list(self, dataset, nmaximgs=-1):
        img_counter, status = 0, 0
        with open(dataset, 'r') as f:
            for line in f:
                if not line.strip():
                    continue
                filename, filename_mask, class_name = line.strip().split(',')
                fp, mp = os.path.abspath(filename), os.path.abspath(filename_mask)
                if not (os.path.isfile(fp) and fp.endswith('.fits')):
                    logger.warning("Image file %s does not exist or has unexpected extension (.fits required)" % filename)
                    status = -1
                    continue
                if not (os.path.isfile(mp) and mp.endswith('.fits')):
                    logger.warning("Mask file %s does not exist or has unexpected extension (.fits required)" % filename_mask)
                    status = -1
                    continue
                if class_name not in self.class_id_map:
                    logger.warning("Image file %s class name (%s) is not present in dictionary, skip it..." % (filename, class_name))
                    status = -1
                    continue
                self._add(fp, [mp], [self.class_id_map.get(class_name)])
                img_counter += 1
                if nmaximgs != -1 and img_counter >= nmaximgs:
                    logger.info("Max number (%d) of desired images reached, stop loading ..." % nmaximgs)
                    break
        if status < 0:
            logger.warning("One or more files have been skipped...")
        if img_counter <= 0:
            logger.error("All files in list have been skipped!")
            return -1
        logger.info("#%d images added in dataset..." % img_counter)
        return 0

    def load_data_from_json_file(self, filename, rootdir='', modify_class_names=True):
        try:
            with open(filename, "r") as jf:
                d = json.load(jf)
        except IOError:
            logger.error("Failed to open file %s, skip it..." % filename)
            return -1
        img_fullpath = os.path.abspath(os.path.join(rootdir, d['img']))
        if not (os.path.isfile(img_fullpath) and img_fullpath.endswith('.fits')):
            logger.warning("Image file %s does not exist or has unexpected extension (.fits required)" % img_fullpath)
            return -1
        metadata = {k: d.get(k) for k in ("telescope", "bkg", "rms", "bmaj", "bmin", "dx", "dy", "nx", "ny")}
        mask_paths, class_ids, near = [], [], []
        for obj in d['objs']:
            mask_fullpath = os.path.abspath(os.path.join(rootdir, obj['mask']))
            if not (os.path.isfile(mask_fullpath) and mask_fullpath.endswith('.fits')):
                logger.error("One or more mask of file %s does not exist or have unexpected extension (.fits required)" % img_fullpath)
                return -1
            class_name = obj['class']
            if modify_class_names:
                if obj.get('nislands', 1) > 1 and class_name == "extended":
                    class_name = 'extended-multisland'
                if obj.get('sidelobe-mixed'):
                    class_name = 'flagged'
                obj['class'] = class_name
            if class_name not in self.class_id_map:
                logger.warning("Image file %s class name (%s) is not present in dictionary, skip it..." % (img_fullpath, class_name))
                continue
            mask_paths.append(mask_fullpath)
            class_ids.append(self.class_id_map.get(class_name))
            near.append(1 if (obj.get('sidelobe-mixed') == 1 or obj.get('sidelobe-near') == 1) else 0)
        self._add(img_fullpath, mask_paths, class_ids, sidelobes_mixed_or_near=near, objs=d['objs'], metadata=metadata)
        return 0

    def load_data_from_json_list(self, filelist, nmaximgs):
        img_counter, status = 0, 0
        with open(filelist, 'r') as f:
            for line in f:
                fn = line.strip()
                if not fn:
                    continue
                if self.load_data_from_json_file(fn, os.path.dirname(fn)) < 0:
                    status = -1
                    continue
                img_counter += 1
                if nmaximgs != -1 and img_counter >= nmaximgs:
                    break
        if status < 0:
            logger.warning("One or more files have been skipped...")
        if img_counter <= 0:
            logger.error("All files in list have been skipped!")
            return -1
        return 0

    def load_gt_masks(self, image_id, binary=True):
        info = self.image_info[image_id]
        mask = None
        for k, filename in enumerate(info["path_masks"]):
            data, _ = utils.read_fits(filename, stretch=False, normalize=False, convertToRGB=False)
            if mask is None:
                mask = np.zeros([data.shape[0], data.shape[1], len(info["path_masks"])], dtype=bool if binary else int)
            mask[:, :, k] = data.astype(bool) if binary else data
        return mask

    def load_mask(self, image_id):
        if self.image_info[image_id]["source"] != "rg-dataset":
            return super(SourceDataset, self).load_mask(image_id)
        mask = self.load_gt_masks(image_id, binary=True)
        return mask, np.full([mask.shape[-1]], self.image_info[image_id]["class_ids"], dtype=np.int32)

    def load_image(self, image_id):
        res = utils.read_fits(self.image_info[image_id]['path'], zscale_contrasts=self.zscale_contrasts)
        if res is None:
            raise IOError("cannot read image " + str(self.image_info[image_id]['path']))
        return res[0]

    def image_reference(self, image_id):
        return self.image_info[image_id]["path"]


def create_train_val_sets_from_filelist(filelist, crossval_size=0.1, train_filename='train.dat', crossval_filename='crossval.dat'):
    """reference: scripts/run.py:821-864 — shuffle the list (Python `random`) and split off the cross-validation part
    (sklearn.model_selection.train_test_split there: ceil(test_size * n) shuffled test samples; here the tail of the
    already shuffled list, which is the same kind of split without the dependency)."""
    import math
    import random
    with open(filelist, 'r') as f:
        data = [ln.strip() for ln in f if ln.strip()]
    if not data:
        logger.error("Given filelist is empty!")
        return []
    if len(data) < 10:
        logger.warning("Given filelist contains less than 10 entries ...")
    random.shuffle(data)
    n_val = max(1, int(math.ceil(float(crossval_size) * len(data)))) if len(data) > 1 else 0
    x_train, x_val = data[:len(data) - n_val], data[len(data) - n_val:]
    for fn, items in ((train_filename, x_train), (crossval_filename, x_val)):
        with open(fn, 'w') as f:
            for item in items:
                f.write("%s\n" % item)
    return [train_filename, crossval_filename]


def create_train_val_datasets(args, train_filename='train.dat', crossval_filename='crossval.dat'):
    """reference: scripts/run.py:893-989"""
    if args.datalist_train and args.datalist_val:
        datalist_train, datalist_val = args.datalist_train, args.datalist_val
    elif args.dataloader in ('datalist', 'datalist_json'):
        lists = create_train_val_sets_from_filelist(args.datalist, args.validation_data_fract, train_filename, crossval_filename)
        if len(lists) != 2:
            return []
        datalist_train, datalist_val = lists
    else:
        logger.error("Invalid/unknown dataloader (%s)!" % args.dataloader)
        return []
    out = []
    for path in (datalist_train, datalist_val):
        ds = SourceDataset()
        ds.set_class_dict(args.classdict)
        ds.zscale_contrasts = [float(x) for x in args.zscale_contrasts.split(',')]
        rc = ds.load_data_from_list(path, args.maxnimgs) if args.dataloader == 'datalist' else ds.load_data_from_json_list(path, args.maxnimgs)
        if rc < 0:
            logger.error("Failed to load dataset from file %s (see logs)..." % path)
            return []
        ds.prepare()
        out.append(ds)
    logger.info("#%d/%d entries in the training/validation sets ..." % (out[0].loaded_imgs, out[1].loaded_imgs))
    return out


def train(args, model, config, datasets):
    """reference: scripts/run.py:1052-1125"""
    if len(datasets) != 2 or datasets[0] is None or datasets[1] is None:
        logger.error("Given dataset list must have size=2!")
        return -1
    logger.info("Training without augmentation steps ...")
    logger.info("Start training ...")
    model.train(datasets[0], datasets[1], learning_rate=config.LEARNING_RATE, epochs=args.nepochs, layers='all',
                n_worker_threads=args.nthreads)
    return 0


def build_parser():
    p = argparse.ArgumentParser(description="Mask R-CNN detect / train on radio maps (B200 build)")
    p.add_argument("command", metavar="<command>", help="'detect', 'test' or 'train'")
    p.add_argument("--dataloader", type=str, default="datalist", help="train: datalist (img,mask,label lines) | datalist_json")
    p.add_argument("--datalist_train", default=None)
    p.add_argument("--datalist_val", default=None)
    p.add_argument("--validation_data_fract", type=float, default=0.1)
    p.add_argument("--nthreads", type=int, default=1)
    p.add_argument("--nepochs", type=int, default=1)
    p.add_argument("--epoch_length", type=int, default=None)
    p.add_argument("--nvalidation_steps", type=int, default=None)
    p.add_argument("--max_gt_instances", type=int, default=300)
    p.add_argument("--rpn_train_anchors_per_image", type=int, default=512)
    p.add_argument("--train_rois_per_image", type=int, default=512)
    for name in ("rpn_class", "rpn_bbox", "mrcnn_class", "mrcnn_bbox", "mrcnn_mask"):      # scripts/run.py:1319-1343
        p.add_argument("--%s_loss_weight" % name, type=float, default=1.0)
        p.add_argument("--%s_loss" % name, dest="%s_loss" % name, action="store_true")
        p.add_argument("--no_%s_loss" % name, dest="%s_loss" % name, action="store_false")
        p.set_defaults(**{"%s_loss" % name: True})
    p.add_argument("--mask_loss_function", type=str, default="binary_crossentropy", choices=["binary_crossentropy", "dice_coef_loss"])
    p.add_argument("--no_augmentation", dest="use_augmentation", action="store_false")
    p.add_argument("--weight_classes", action="store_true")
    p.add_argument("--imgsize", dest="imgsize", type=int, default=256)
    p.add_argument("--grayimg", dest="grayimg", action="store_true")
    p.add_argument("--no_uint8", dest="to_uint8", action="store_false")
    p.add_argument("--no_zscale", dest="zscale", action="store_false")
    p.add_argument("--zscale_contrasts", dest="zscale_contrasts", type=str, default="0.25,0.25,0.25")
    p.add_argument("--biascontrast", dest="biascontrast", action="store_true")
    p.add_argument("--bias", type=float, default=0.5, help="bias of the bias-contrast stretch (only with --biascontrast, which this build rejects)")
    p.add_argument("--contrast", type=float, default=1.0, help="contrast of the bias-contrast stretch (only with --biascontrast)")
    # accepted for command lines written for the reference (scripts/run.py:1289-1297, 1360-1370); see validate_args
    p.add_argument("--remap_classids", dest="remap_classids", action="store_true")
    p.add_argument("--classid_remap_dict", type=str, default="")
    p.add_argument("--datadir", required=False, default=None)
    p.add_argument("--consider_sources_near_mixed_sidelobes", dest="consider_sources_near_mixed_sidelobes", action="store_true")
    p.add_argument("--no_consider_sources_near_mixed_sidelobes", dest="consider_sources_near_mixed_sidelobes", action="store_false")
    p.set_defaults(consider_sources_near_mixed_sidelobes=True)
    p.add_argument("--detect_outfile", type=str, default="", help="output plot PNG of the reference: accepted, no plot is drawn")
    p.add_argument("--no_norm_img", dest="norm_img", action="store_false")
    p.add_argument("--classdict", dest="classdict", type=str, default='{"sidelobe":1,"source":2,"galaxy":3}')
    p.add_argument("--classdict_model", dest="classdict_model", type=str, default="")
    p.add_argument("--datalist", required=False)
    p.add_argument("--maxnimgs", type=int, default=-1)
    p.add_argument("--weights", required=False)
    p.add_argument("--random_weights", type=int, default=None, metavar="SEED",
                   help="extension: seeded random weights instead of --weights (the shipped .h5 is a Git-LFS pointer)")
    p.add_argument("--logs", default=os.path.join(os.getcwd(), "logs"))
    p.add_argument("--ngpu", type=int, default=1)
    p.add_argument("--nimg_per_gpu", type=int, default=1)
    p.add_argument("--rpn_anchor_scales", type=str, default="4,8,16,32,64")
    p.add_argument("--backbone", type=str, default="resnet101")
    p.add_argument("--backbone_strides", type=str, default="4,8,16,32,64")
    p.add_argument("--rpn_nms_threshold", type=float, default=0.7)
    p.add_argument("--rpn_anchor_ratios", type=str, default="0.5,1,2")
    p.add_argument("--exclude_first_layer_weights", action="store_true")
    p.add_argument("--scoreThr", type=float, default=0.7)
    p.add_argument("--iouThr", type=float, default=0.6)
    p.add_argument("--image", type=str)
    p.add_argument("--xmin", type=int, default=-1)
    p.add_argument("--xmax", type=int, default=-1)
    p.add_argument("--ymin", type=int, default=-1)
    p.add_argument("--ymax", type=int, default=-1)
    p.add_argument("--detect_outfile_json", type=str, default="")
    p.add_argument("--split_img_in_tiles", action="store_true")
    p.add_argument("--tile_xsize", type=int, default=512, help="Sub image size in pixel along x")
    p.add_argument("--tile_ysize", type=int, default=512, help="Sub image size in pixel along y")
    p.add_argument("--tile_xstep", type=float, default=1.0, help="Sub image step fraction along x (=1 means no overlap)")
    p.add_argument("--tile_ystep", type=float, default=1.0, help="Sub image step fraction along y (=1 means no overlap)")
    return p


def parse_args(argv=None):
    return build_parser().parse_args(argv)


def validate_args(args):
    """scripts/run.py:1387-1443 for the commands kept here."""
    if args.command not in ("detect", "test", "train"):
        logger.error("Command '%s' is not available in the B200 build (detect / test / train)." % args.command)
        return -1
    if args.command == "train":
        has_split = args.datalist_train and args.datalist_val
        if not has_split and not (args.datalist and os.path.isfile(args.datalist)):
            logger.error("Argument --datalist (or --datalist_train and --datalist_val) is required for train task!")
            return -1
        if args.weight_classes:
            logger.error("Option --weight_classes is not supported by the B200 build")
            return -1
        args.use_augmentation = False        # imgaug is not part of this build: training runs without augmentation
    if args.command == "detect":
        if not args.image:
            logger.error("Argument --image is required for detect task!")
            return -1
        if not os.path.isfile(args.image) or not args.image.endswith(".fits"):
            logger.error("Image %s does not exist or has not .fits extension!" % args.image)
            return -1
    if args.command == "test" and not (args.datalist and os.path.isfile(args.datalist)):
        logger.error("Argument --datalist (existing file) is required for test task!")
        return -1
    if args.command != "train" and not args.weights and args.random_weights is None:
        logger.error("Argument --weights is required (or --random_weights SEED)")
        return -1
    for flag, ok in (("--grayimg", not args.grayimg), ("--no_uint8", args.to_uint8), ("--no_zscale", args.zscale),
                     ("--biascontrast", not args.biascontrast), ("--no_norm_img", args.norm_img),
                     ("--backbone != resnet101", args.backbone == "resnet101"),
                     ("--ngpu > 1 (run one process per GPU instead)", args.ngpu == 1)):
        if not ok:
            logger.error("Option %s is not supported by the B200 build (SURVEY.md §8a row a17)" % flag)
            return -1
    if args.remap_classids:
        if args.classid_remap_dict == "":          # scripts/run.py:1438-1441
            logger.error("Classid remap dictionary is empty (you need to provide one if you give the option --remap_classids)!")
            return -1
        logger.error("Option --remap_classids belongs to the ground-truth metrics of the reference (ModelTester), which the B200 "
                     "build does not provide (SURVEY.md §8: out of scope)")
        return -1
    if args.dataloader in ("datadir", "datadir_json"):
        logger.error("Data loader '%s' (--datadir tree search) is not supported by the B200 build: use datalist / datalist_json"
                     % args.dataloader)
        return -1
    if args.command == "train" and args.mask_loss_function != "binary_crossentropy":
        logger.error("Option --mask_loss_function %s is not supported by the B200 build (binary_crossentropy only)" % args.mask_loss_function)
        return -1
    if args.detect_outfile:
        logger.warning("--detect_outfile %s: plots are not drawn by the B200 build (DESIGN.md §6), option ignored" % args.detect_outfile)
    try:
        nclasses = len(json.loads(args.classdict_model or args.classdict)) + 1
    except Exception:
        logger.error("Cannot parse --classdict / --classdict_model as a JSON dictionary")
        return -1
    if nclasses > 6:
        logger.error("%d classes (+ background) given: the B200 build supports at most 6 classes including background" % (nclasses - 1))
        return -1
    return 0


def make_config(args):
    classdict = json.loads(args.classdict_model or args.classdict)
    config = SDetectorConfig()
    config.GPU_COUNT = 1
    config.IMAGES_PER_GPU = args.nimg_per_gpu
    config.BATCH_SIZE = args.nimg_per_gpu
    config.NUM_CLASSES = len(classdict) + 1
    config.CLASS_NAMES = ["bkg"] + [k for k, _ in sorted(classdict.items(), key=lambda kv: kv[1])]
    config.IMAGE_META_SIZE = 1 + 3 + 3 + 4 + 1 + config.NUM_CLASSES
    config.RPN_ANCHOR_SCALES = tuple(int(x) for x in args.rpn_anchor_scales.split(","))
    config.BACKBONE = args.backbone
    config.BACKBONE_STRIDES = [int(x) for x in args.backbone_strides.split(",")]
    config.RPN_NMS_THRESHOLD = args.rpn_nms_threshold
    config.RPN_ANCHOR_RATIOS = [float(x) for x in args.rpn_anchor_ratios.split(",")]
    config.IMAGE_MIN_DIM = config.IMAGE_MAX_DIM = args.imgsize
    config.IMAGE_SHAPE = np.array([args.imgsize, args.imgsize, config.IMAGE_CHANNEL_COUNT])
    config.ZSCALE_CONTRASTS = [float(x) for x in args.zscale_contrasts.split(",")]
    config.IOU_THR = args.iouThr
    config.SCORE_THR = args.scoreThr
    config.IMG_PATH = args.image
    config.IMG_XMIN, config.IMG_XMAX, config.IMG_YMIN, config.IMG_YMAX = args.xmin, args.xmax, args.ymin, args.ymax
    config.SPLIT_IMG_IN_TILES = args.split_img_in_tiles
    config.TILE_XSIZE, config.TILE_YSIZE = args.tile_xsize, args.tile_ysize
    config.TILE_XSTEP, config.TILE_YSTEP = args.tile_xstep, args.tile_ystep
    config.MAX_NTASKS_PER_WORKER = 100000
    config.OUTFILE_JSON = args.detect_outfile_json
    if args.command == "train":          # scripts/run.py:1627-1672 (training values of SDetectorConfig + CLI overrides)
        config.MAX_GT_INSTANCES = args.max_gt_instances
        config.RPN_TRAIN_ANCHORS_PER_IMAGE = args.rpn_train_anchors_per_image
        config.TRAIN_ROIS_PER_IMAGE = args.train_rois_per_image
        config.LEARNING_RATE = 0.0005
        config.USE_MINI_MASK = False
        config.LOSS_WEIGHTS = {"%s_loss" % n: getattr(args, "%s_loss_weight" % n) for n in
                               ("rpn_class", "rpn_bbox", "mrcnn_class", "mrcnn_bbox", "mrcnn_mask")}
        config.USE_LOSSES = {"%s_loss" % n: getattr(args, "%s_loss" % n) for n in
                             ("rpn_class", "rpn_bbox", "mrcnn_class", "mrcnn_bbox", "mrcnn_mask")}
        config.MASK_LOSS_FUNCTION = args.mask_loss_function
    return config


def load_model(args, config):
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank)      # one process per GPU: everything of this rank (read_fits included) on its GPU
    mode = "training" if args.command == "train" else "inference"
    model = modellib.MaskRCNN(mode=mode, config=config, model_dir=args.logs, device=local_rank)
    if mode == "training" and not args.weights and args.random_weights is None:
        args.random_weights = 0                  # training from scratch: seeded random initialisation
    if args.random_weights is not None:
        sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
        import synth
        model.set_weights(synth.make_random_weights(args.random_weights, config.NUM_CLASSES))
    elif args.exclude_first_layer_weights:
        model.load_weights(args.weights, by_name=True, exclude="conv1")
    else:
        model.load_weights(args.weights, by_name=True)
    return model


def result_to_json(name, r, config):
    objs = []
    for i in range(len(r["class_ids"])):
        if r["scores"][i] < config.SCORE_THR:
            continue
        y1, x1, y2, x2 = [int(v) for v in r["rois"][i]]
        cid = int(r["class_ids"][i])
        objs.append({"name": "S%d" % (len(objs) + 1), "x1": x1, "x2": x2, "y1": y1, "y2": y2, "class_id": cid,
                     "class_name": config.CLASS_NAMES[cid] if cid < len(config.CLASS_NAMES) else str(cid),
                     "score": float(r["scores"][i]), "npix": int(r["masks"][:, :, i].sum())})
    return {"image": name, "objs": objs, "ndet_raw": int(len(r["class_ids"]))}


def detect(args, model, config):
    if args.split_img_in_tiles:
        # scripts/run.py:1176-1183: SFinder.run_parallel(); the workers are the torch.distributed ranks (one per GPU)
        if int(os.environ.get("WORLD_SIZE", "1")) > 1:
            import torch.distributed as dist
            if not dist.is_initialized():
                dist.init_process_group("gloo")          # catalogues only (all_gather_object); no data-path collective
        sfinder = SFinder(model, config)
        sfinder.outfile_json = args.detect_outfile_json
        if sfinder.run_parallel() < 0:
            logger.error("sfinder run failed, see logs...")
            return -1
        return 0
    res = utils.read_fits(args.image, args.xmin, args.xmax, args.ymin, args.ymax, zscale_contrasts=config.ZSCALE_CONTRASTS)
    if res is None:
        logger.error("Failed to read image %s!" % args.image)
        return -1
    image, header = res
    image_id = os.path.splitext(os.path.basename(args.image))[0]
    analyzer = Analyzer(_FirstOfBatch(model, config.BATCH_SIZE), config)
    analyzer.draw = False
    analyzer.write_to_ds9 = False
    analyzer.write_to_json = True
    analyzer.outfile_json = args.detect_outfile_json
    analyzer.iou_thr = config.IOU_THR
    analyzer.score_thr = config.SCORE_THR
    if analyzer.predict(image, image_id, header=header) < 0:
        logger.error("Failed to run model prediction on image %s!" % args.image)
        return -1
    if not analyzer.bboxes:
        logger.info("No object detected in image %s ..." % args.image)
        return 0
    logger.info("#%d objects found in image %s ..." % (len(analyzer.bboxes), args.image))
    return 0


class _FirstOfBatch:
    """detect([image]) for a model built with BATCH_SIZE > 1: the image fills the batch, the first result is returned."""

    def __init__(self, model, batch):
        self._model, self._batch = model, batch
        self._device, self._stream = model._device, model._stream

    def detect(self, images, verbose=0):
        return self._model.detect(list(images) * self._batch, verbose=verbose)[:1]


def test(args, model, config):
    files = [ln.split(",")[0].strip() for ln in open(args.datalist) if ln.strip() and not ln.startswith("#")]
    if args.maxnimgs > 0:
        files = files[:args.maxnimgs]
    B, outs = config.BATCH_SIZE, []
    for i in range(0, len(files), B):
        batch, names = [], files[i:i + B]
        for fn in names:
            res = utils.read_fits(fn, zscale_contrasts=config.ZSCALE_CONTRASTS)
            if res is None:
                logger.error("Failed to read image %s!" % fn)
                return -1
            batch.append(res[0])
        pad = B - len(batch)
        results = model.detect(batch + [batch[-1]] * pad)
        outs += [result_to_json(os.path.basename(n), r, config) for n, r in zip(names, results)]
    path = args.detect_outfile_json or "out_test.json"
    with open(path, "w") as f:
        json.dump(outs, f, indent=1)
    logger.info("detections of %d images written to %s" % (len(outs), path))
    return 0


def main(argv=None):
    try:
        args = parse_args(argv)
    except SystemExit:
        return 1
    if validate_args(args) < 0:
        logger.error("Argument validation failed, exit ...")
        return 1
    config = make_config(args)
    try:
        if args.command == "train":
            world = int(os.environ.get("WORLD_SIZE", "1"))
            if world > 1:
                import torch
                import torch.distributed as dist
                torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
                if not dist.is_initialized():
                    dist.init_process_group("nccl")
            datasets = create_train_val_datasets(args, "train_%d.dat" % os.getpid(), "crossval_%d.dat" % os.getpid())
            if not datasets:
                logger.error("Failed to create train/validation datasets!")
                return 1
            if args.epoch_length is not None:
                config.STEPS_PER_EPOCH = args.epoch_length
            else:
                config.STEPS_PER_EPOCH = max(1, datasets[0].loaded_imgs // (config.BATCH_SIZE * world))
            config.VALIDATION_STEPS = args.nvalidation_steps if args.nvalidation_steps is not None else \
                max(1, datasets[1].loaded_imgs // (config.BATCH_SIZE * world))
            model = load_model(args, config)
            status = train(args, model, config, datasets)
            return 0 if status == 0 else 1
        model = load_model(args, config)
        status = detect(args, model, config) if args.command == "detect" else test(args, model, config)
    except Exception as e:      # noqa: BLE001 — the reference's main() turns failures into exit code 1
        logging.getLogger("mrcnn").error("%s failed: %s" % (args.command, e))
        return 1
    return 0 if status == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
