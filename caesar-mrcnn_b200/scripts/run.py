#!/usr/bin/env python
"""run.py — detect/test CLI of the B200 build, flag-compatible with the detect-path subset of the
reference's scripts/run.py (parse_args :1263-1384, InferenceConfig :1652-1706, detect :1172-1189).

  run.py detect --image data/galaxy0002.fits --weights share/mrcnn_weights.h5 [--imgsize 256 ...]
  run.py test   --datalist images.txt        --weights share/mrcnn_weights.h5

`detect` reads the FITS image (zscale + uint8 RGB on the GPU) and hands it to `Analyzer.predict` exactly as the
reference's SFinder.run does (mrcnn/sfinder.py:485-493): MaskRCNN.detect, then extract_det_masks (score filter,
merging of connected same-class masks, best-of-overlapping selection) on the GPU, then the reference's JSON catalogue
(`out_<image>.json` or --detect_outfile_json; keys name,x1,x2,y1,y2,class_id,class_name,score,pixels,vertexes,edge).
`test` writes the raw detections with score >= --scoreThr of every FITS file of a list (batched by --nimg_per_gpu).
With --split_img_in_tiles the image is cut into --tile_xsize x --tile_ysize tiles (SFinder.run_parallel,
mrcnn/sfinder.py:549-640): the tiles go through the detector in batches of --nimg_per_gpu, sources cut by tile borders
are merged, and `catalog_<image>.json` is written; launched under torchrun, every rank (GPU) takes the tiles the
reference's MPI ranks would. PNG / DS9 output, the ground-truth metrics of `test`, source parameters (WCS / flux) and
`train` are outside the path rebuilt here and are rejected or skipped.
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


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="Mask R-CNN detect on radio maps (B200 build)")
    p.add_argument("command", metavar="<command>", help="'detect' or 'test' ('train' is not part of this build)")
    p.add_argument("--imgsize", dest="imgsize", type=int, default=256)
    p.add_argument("--grayimg", dest="grayimg", action="store_true")
    p.add_argument("--no_uint8", dest="to_uint8", action="store_false")
    p.add_argument("--no_zscale", dest="zscale", action="store_false")
    p.add_argument("--zscale_contrasts", dest="zscale_contrasts", type=str, default="0.25,0.25,0.25")
    p.add_argument("--biascontrast", dest="biascontrast", action="store_true")
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
    return p.parse_args(argv)


def validate_args(args):
    """scripts/run.py:1387-1443 for the commands kept here."""
    if args.command not in ("detect", "test"):
        logger.error("Command '%s' is not available in the B200 build (only detect/test)." % args.command)
        return -1
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
    if not args.weights and args.random_weights is None:
        logger.error("Argument --weights is required (or --random_weights SEED)")
        return -1
    for flag, ok in (("--grayimg", not args.grayimg), ("--no_uint8", args.to_uint8), ("--no_zscale", args.zscale),
                     ("--biascontrast", not args.biascontrast), ("--no_norm_img", args.norm_img),
                     ("--backbone != resnet101", args.backbone == "resnet101"),
                     ("--ngpu > 1 (run one process per GPU instead)", args.ngpu == 1)):
        if not ok:
            logger.error("Option %s is not supported by the B200 build (SURVEY.md §8a row a17)" % flag)
            return -1
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
    return config


def load_model(args, config):
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank)      # one process per GPU: everything of this rank (read_fits included) on its GPU
    model = modellib.MaskRCNN(mode="inference", config=config, model_dir=args.logs, device=local_rank)
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
        model = load_model(args, config)
        status = detect(args, model, config) if args.command == "detect" else test(args, model, config)
    except Exception as e:      # noqa: BLE001 — the reference's main() turns failures into exit code 1
        logging.getLogger("mrcnn").error("%s failed: %s" % (args.command, e))
        return 1
    return 0 if status == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
