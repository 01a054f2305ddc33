"""Train mode of the B200 build (BASELINE.json configs[4]; SURVEY.md §8e training, §8f rank 4).

Reference                                             here
  data_generator / load_image_gt / build_rpn_targets  host numpy, same numpy-RNG draw order   (mrcnn/model.py:1277-1381,
                                                                                               1536-1644, 1721-1904)
  training graph                                      TrainGraph.forward_backward               (mrcnn/model.py:2003-2132)
  DetectionTargetLayer                                csrc/train_ops.cu detection_targets_kernel (mrcnn/model.py:570-763)
  PyramidROIAlign (+ gradient)                        csrc/roialign.cu + roialign_backward_kernel
  ProposalLayer                                       csrc/proposal.cu (POST_NMS_ROIS_TRAINING rows)
  five losses                                         losses()                                   (mrcnn/model.py:1098-1270)
  keras SGD + clipnorm + L2 regulariser               csrc/train_ops.cu sgd_* over one flat buffer (mrcnn/model.py:2259-2297)
  ParallelModel (towers + CPU merge)                  one process per GPU, GradReducer: bucketed NCCL all-reduce of the
                                                      flat gradient buffer overlapped with backward (mrcnn/parallel_model.py)

Parameters live in ONE flat float32 buffer (+ gradient, momentum and bf16 operand copies of the same shape): a bucket of
the gradient all-reduce is a slice of it, the optimiser is two kernel launches, and the bf16 copy the next forward
multiplies with is written by the optimiser itself.  Convolution kernels are stored as [Cout, KH, KW, Cin] — the K-major
B operand of the implicit GEMM and, seen through a permuted view, a channels-last OIHW tensor.  The dense contractions
of the training graph go through torch's bf16 convolution / matmul on NHWC tensors (library calls); everything else on
the step is this package's CUDA.  There is no CPU path: the graph needs a CUDA device.
"""
import ctypes
import logging
import re

import numpy as np

from . import _native, utils

BN_EPS = 1e-3
LOSS_NAMES = ["rpn_class_loss", "rpn_bbox_loss", "mrcnn_class_loss", "mrcnn_bbox_loss", "mrcnn_mask_loss"]


# ============================================================================================================
# host data path
# ============================================================================================================

def build_rpn_targets(image_shape, anchors, gt_class_ids, gt_boxes, config):
    """reference: mrcnn/model.py:1536-1644.  -> rpn_match [A] int32 (1 positive, -1 negative, 0 neutral),
    rpn_bbox [RPN_TRAIN_ANCHORS_PER_IMAGE, 4] float64 deltas of the positive anchors in anchor order (zero padded).
    The two random sub-samplings draw from numpy's global generator with the reference's calls in the reference's order
    (np.random.choice(ids, extra, replace=False)), so a seeded run selects the same anchors."""
    anchors = np.asarray(anchors)
    gt_class_ids = np.asarray(gt_class_ids)
    gt_boxes = np.asarray(gt_boxes)
    rpn_match = np.zeros([anchors.shape[0]], dtype=np.int32)
    rpn_bbox = np.zeros((config.RPN_TRAIN_ANCHORS_PER_IMAGE, 4))
    crowd_ix = np.where(gt_class_ids < 0)[0]
    if crowd_ix.shape[0] > 0:
        keep = np.where(gt_class_ids > 0)[0]
        crowd_boxes = gt_boxes[crowd_ix]
        gt_class_ids, gt_boxes = gt_class_ids[keep], gt_boxes[keep]
        no_crowd = np.amax(utils.compute_overlaps(anchors, crowd_boxes), axis=1) < 0.001
    else:
        no_crowd = np.ones([anchors.shape[0]], dtype=bool)
    overlaps = utils.compute_overlaps(anchors, gt_boxes)
    arg = np.argmax(overlaps, axis=1)
    best = overlaps[np.arange(overlaps.shape[0]), arg]
    rpn_match[(best < 0.3) & no_crowd] = -1
    rpn_match[np.argwhere(overlaps == np.max(overlaps, axis=0))[:, 0]] = 1      # every GT box keeps its best anchor(s)
    rpn_match[best >= 0.7] = 1
    ids = np.where(rpn_match == 1)[0]
    extra = len(ids) - (config.RPN_TRAIN_ANCHORS_PER_IMAGE // 2)
    if extra > 0:
        rpn_match[np.random.choice(ids, extra, replace=False)] = 0
    ids = np.where(rpn_match == -1)[0]
    extra = len(ids) - (config.RPN_TRAIN_ANCHORS_PER_IMAGE - np.sum(rpn_match == 1))
    if extra > 0:
        rpn_match[np.random.choice(ids, extra, replace=False)] = 0
    ids = np.where(rpn_match == 1)[0]
    if len(ids):
        a = anchors[ids]
        gt = gt_boxes[arg[ids]]
        gt_h, gt_w = gt[:, 2] - gt[:, 0], gt[:, 3] - gt[:, 1]
        gt_cy, gt_cx = gt[:, 0] + 0.5 * gt_h, gt[:, 1] + 0.5 * gt_w
        a_h, a_w = a[:, 2] - a[:, 0], a[:, 3] - a[:, 1]
        a_cy, a_cx = a[:, 0] + 0.5 * a_h, a[:, 1] + 0.5 * a_w
        d = np.stack([(gt_cy - a_cy) / a_h, (gt_cx - a_cx) / a_w, np.log(gt_h / a_h), np.log(gt_w / a_w)], axis=1)
        rpn_bbox[:len(ids)] = d / config.RPN_BBOX_STD_DEV
    return rpn_match, rpn_bbox


def load_image_gt(dataset, config, image_id, augment=False, augmentation=None, use_mini_mask=False):
    """reference: mrcnn/model.py:1277-1381 -> image, image_meta, class_ids, bbox, mask.  imgaug pipelines are not
    supported in this build (`augmentation` must be None); the deprecated `augment` flag flips left/right with Python's
    `random.randint(0, 1)` as the reference does."""
    from .model import compose_image_meta
    if augmentation is not None:
        raise NotImplementedError("mrcnn (B200 build): imgaug augmentation pipelines are not available")
    image = dataset.load_image(image_id)
    mask, class_ids = dataset.load_mask(image_id)
    original_shape = image.shape
    image, window, scale, padding, crop = utils.resize_image(image, min_dim=config.IMAGE_MIN_DIM, min_scale=config.IMAGE_MIN_SCALE,
                                                             max_dim=config.IMAGE_MAX_DIM, mode=config.IMAGE_RESIZE_MODE)
    mask = utils.resize_mask(mask, scale, padding, crop)
    if augment:
        import random
        logging.warning("'augment' is deprecated. Use 'augmentation' instead.")
        if random.randint(0, 1):
            image, mask = np.fliplr(image), np.fliplr(mask)
    keep = np.sum(mask, axis=(0, 1)) > 0                    # instances that were cropped away
    mask = mask[:, :, keep]
    class_ids = class_ids[keep]
    bbox = utils.extract_bboxes(mask)
    active_class_ids = np.zeros([dataset.num_classes], dtype=np.int32)
    active_class_ids[dataset.source_class_ids[dataset.image_info[image_id]["source"]]] = 1
    if use_mini_mask:
        mask = utils.minimize_mask(bbox, mask, config.MINI_MASK_SHAPE)
    image_meta = compose_image_meta(image_id, original_shape, image.shape, window, scale, active_class_ids)
    return image, image_meta, class_ids, bbox, mask


def data_generator(dataset, config, shuffle=True, augment=False, augmentation=None, random_rois=0, batch_size=1,
                   detection_targets=False, no_augmentation_sources=None):
    """reference: mrcnn/model.py:1721-1904.  Yields (inputs, outputs) with inputs = [images, image_meta, rpn_match,
    rpn_bbox, gt_class_ids, gt_boxes, gt_masks] (same shapes and dtypes) and outputs = [].  `random_rois` /
    `detection_targets` (debug modes of the reference) are not implemented."""
    from .model import mold_image
    if random_rois or detection_targets:
        raise NotImplementedError("mrcnn (B200 build): data_generator(random_rois / detection_targets) debug modes")
    b, image_index, error_count = 0, -1, 0
    image_ids = np.copy(dataset.image_ids)
    no_augmentation_sources = no_augmentation_sources or []
    anchors = utils.generate_pyramid_anchors(config.RPN_ANCHOR_SCALES, config.RPN_ANCHOR_RATIOS,
                                             utils.compute_backbone_shapes(config, config.IMAGE_SHAPE),
                                             config.BACKBONE_STRIDES, config.RPN_ANCHOR_STRIDE)
    while True:
        try:
            image_index = (image_index + 1) % len(image_ids)
            if shuffle and image_index == 0:
                np.random.shuffle(image_ids)
            image_id = image_ids[image_index]
            aug = None if dataset.image_info[image_id]['source'] in no_augmentation_sources else augmentation
            image, image_meta, gt_class_ids, gt_boxes, gt_masks = load_image_gt(
                dataset, config, image_id, augment=augment, augmentation=aug, use_mini_mask=config.USE_MINI_MASK)
            if not np.any(gt_class_ids > 0):
                continue
            rpn_match, rpn_bbox = build_rpn_targets(image.shape, anchors, gt_class_ids, gt_boxes, config)
            if b == 0:
                batch_image_meta = np.zeros((batch_size,) + image_meta.shape, dtype=image_meta.dtype)
                batch_rpn_match = np.zeros([batch_size, anchors.shape[0], 1], dtype=rpn_match.dtype)
                batch_rpn_bbox = np.zeros([batch_size, config.RPN_TRAIN_ANCHORS_PER_IMAGE, 4], dtype=rpn_bbox.dtype)
                batch_images = np.zeros((batch_size,) + image.shape, dtype=np.float32)
                batch_gt_class_ids = np.zeros((batch_size, config.MAX_GT_INSTANCES), dtype=np.int32)
                batch_gt_boxes = np.zeros((batch_size, config.MAX_GT_INSTANCES, 4), dtype=np.int32)
                batch_gt_masks = np.zeros((batch_size, gt_masks.shape[0], gt_masks.shape[1], config.MAX_GT_INSTANCES),
                                          dtype=gt_masks.dtype)
            if gt_boxes.shape[0] > config.MAX_GT_INSTANCES:
                ids = np.random.choice(np.arange(gt_boxes.shape[0]), config.MAX_GT_INSTANCES, replace=False)
                gt_class_ids, gt_boxes, gt_masks = gt_class_ids[ids], gt_boxes[ids], gt_masks[:, :, ids]
            batch_image_meta[b] = image_meta
            batch_rpn_match[b] = rpn_match[:, np.newaxis]
            batch_rpn_bbox[b] = rpn_bbox
            batch_images[b] = mold_image(image.astype(np.float32), config)
            batch_gt_class_ids[b, :gt_class_ids.shape[0]] = gt_class_ids
            batch_gt_boxes[b, :gt_boxes.shape[0]] = gt_boxes
            batch_gt_masks[b, :, :, :gt_masks.shape[-1]] = gt_masks
            b += 1
            if b >= batch_size:
                yield [batch_images, batch_image_meta, batch_rpn_match, batch_rpn_bbox, batch_gt_class_ids, batch_gt_boxes,
                       batch_gt_masks], []
                b = 0
        except (GeneratorExit, KeyboardInterrupt):
            raise
        except Exception:
            logging.exception("Error processing image {}".format(dataset.image_info[image_id]))
            error_count += 1
            if error_count > 5:
                raise


# ============================================================================================================
# parameters
# ============================================================================================================

def layer_specs(config):
    """Weighted layers of the training graph in forward order: (name, kind, keras kernel shape), kind in
    conv | bn | dense | deconv (same inventory as the inference engine's layer table)."""
    nc, fc, pyr = int(config.NUM_CLASSES), int(config.FPN_CLASSIF_FC_LAYERS_SIZE), int(config.TOP_DOWN_PYRAMID_SIZE)
    apl = len(config.RPN_ANCHOR_RATIOS)
    specs = [("conv1", "conv", (7, 7, 3, 64)), ("bn_conv1", "bn", (64,))]
    cin = 64
    for stage, nblocks, (f1, f2, f3) in ((2, 3, (64, 64, 256)), (3, 4, (128, 128, 512)), (4, 23, (256, 256, 1024)),
                                         (5, 3, (512, 512, 2048))):
        for i in range(nblocks):
            blk = chr(97 + i)
            base, bnb = "res%d%s_branch" % (stage, blk), "bn%d%s_branch" % (stage, blk)
            specs += [(base + "2a", "conv", (1, 1, cin, f1)), (bnb + "2a", "bn", (f1,)),
                      (base + "2b", "conv", (3, 3, f1, f2)), (bnb + "2b", "bn", (f2,)),
                      (base + "2c", "conv", (1, 1, f2, f3)), (bnb + "2c", "bn", (f3,))]
            if i == 0:
                specs += [(base + "1", "conv", (1, 1, cin, f3)), (bnb + "1", "bn", (f3,))]
            cin = f3
    specs += [("fpn_c5p5", "conv", (1, 1, 2048, pyr)), ("fpn_c4p4", "conv", (1, 1, 1024, pyr)),
              ("fpn_c3p3", "conv", (1, 1, 512, pyr)), ("fpn_c2p2", "conv", (1, 1, 256, pyr))]
    specs += [(n, "conv", (3, 3, pyr, pyr)) for n in ("fpn_p2", "fpn_p3", "fpn_p4", "fpn_p5")]
    specs += [("rpn_conv_shared", "conv", (3, 3, pyr, 512)), ("rpn_class_raw", "conv", (1, 1, 512, 2 * apl)),
              ("rpn_bbox_pred", "conv", (1, 1, 512, 4 * apl))]
    ps = int(config.POOL_SIZE)
    specs += [("mrcnn_class_conv1", "conv", (ps, ps, pyr, fc)), ("mrcnn_class_bn1", "bn", (fc,)),
              ("mrcnn_class_conv2", "conv", (1, 1, fc, fc)), ("mrcnn_class_bn2", "bn", (fc,)),
              ("mrcnn_class_logits", "dense", (fc, nc)), ("mrcnn_bbox_fc", "dense", (fc, 4 * nc))]
    for i in range(1, 5):
        specs += [("mrcnn_mask_conv%d" % i, "conv", (3, 3, pyr, pyr)), ("mrcnn_mask_bn%d" % i, "bn", (pyr,))]
    specs += [("mrcnn_mask_deconv", "deconv", (2, 2, pyr, pyr)), ("mrcnn_mask", "conv", (1, 1, pyr, nc))]
    return specs


def _round8(n):
    return (int(n) + 7) // 8 * 8


class ParamStore(object):
    """Flat parameter storage.  Layout of `w` (float32; `g` gradient, `v` momentum, `wb` bf16 copy share it):
        [ biases of the convs followed by a BatchNorm | BN gammas | BN betas | other biases | kernels in forward layer order ]
    every tensor starting at a multiple of 8 elements.  The first three groups (and the moving means / variances in
    `stats`, never trained: TRAIN_BN=False, mrcnn/config.py:216) list the BatchNorm layers in the same order with the same
    offsets, so the folded affine of EVERY layer (scale = gamma / sqrt(var + eps), shift = (bias - mean) * scale + beta)
    and its backward are a handful of vector operations over whole groups.  Gradients of the kernels become available
    in reverse layer order, so all-reduce buckets are cut from the END of the buffer; the small vectors at the front
    complete last.  kernel storage: conv [Cout,KH,KW,Cin], dense [out,in], deconv [KH,KW,Cout,Cin] (its Keras layout)."""

    def __init__(self, config, torch, device):
        self.torch, self.device = torch, device
        self.specs = layer_specs(config)
        self.entries = {}                    # (layer, role) -> (offset, shape) in w;  role: kernel bias gamma beta
        self.stat_entries = {}               # (layer, role) -> (offset, shape) in stats;  role: mean var
        self.bn_of = {}                      # conv layer -> its BatchNorm layer
        self.bn_rel = {}                     # BatchNorm layer -> (offset inside a BN group, channels)
        rel = 0
        for i, (name, kind, shape) in enumerate(self.specs):
            if kind == "bn":
                self.bn_of[self.specs[i - 1][0]] = name
                self.bn_rel[name] = (rel, int(shape[0]))
                rel += _round8(shape[0])
        self.bn_elements = rel
        off = 0
        order = []
        self.group = {}

        def add(name, role, shp):
            nonlocal off
            self.entries[(name, role)] = (off, shp)
            order.append((name, role, off, int(np.prod(shp))))
            off += _round8(np.prod(shp))

        def cout_of(kind, shape):
            return shape[2] if kind == "deconv" else shape[-1]

        self.group["bias_bn"] = off
        for name, kind, shape in self.specs:
            if kind != "bn" and name in self.bn_of:
                add(name, "bias", (cout_of(kind, shape),))
        for role in ("gamma", "beta"):
            self.group[role] = off
            for name, kind, shape in self.specs:
                if kind == "bn":
                    add(name, role, (shape[0],))
        for name, kind, shape in self.specs:
            if kind != "bn" and name not in self.bn_of:
                add(name, "bias", (cout_of(kind, shape),))
        for name, kind, shape in self.specs:
            if kind == "conv":
                add(name, "kernel", (shape[3], shape[0], shape[1], shape[2]))
            elif kind == "dense":
                add(name, "kernel", (shape[1], shape[0]))
            elif kind == "deconv":
                add(name, "kernel", tuple(shape))
        self.order = order
        self.n = off
        for k, role in enumerate(("mean", "var")):
            for name, (r, c) in self.bn_rel.items():
                self.stat_entries[(name, role)] = (k * self.bn_elements + r, (c,))
        self.w = torch.zeros(self.n, dtype=torch.float32, device=device)
        self.g = torch.zeros(self.n, dtype=torch.float32, device=device)
        self.v = torch.zeros(self.n, dtype=torch.float32, device=device)
        self.wb = torch.zeros(self.n, dtype=torch.bfloat16, device=device)
        self.stats = torch.zeros(max(2 * self.bn_elements, 8), dtype=torch.float32, device=device)
        self.kinds = {name: kind for name, kind, _ in self.specs}
        self.keras_shapes = {name: shape for name, _, shape in self.specs}
        self.trainable_elements = sum(sz for _, _, _, sz in order)

    def bn_group(self, buf, role):
        """whole-group view (all BatchNorm layers) of `role` in bias_bn | gamma | beta (buf = w, g, ...) or mean | var (stats)"""
        if role in ("mean", "var"):
            o = 0 if role == "mean" else self.bn_elements
            return self.stats[o:o + self.bn_elements]
        o = self.group[role]
        return buf[o:o + self.bn_elements]

    # -- views ------------------------------------------------------------------------------------------------
    def view(self, buf, name, role):
        off, shp = self.entries[(name, role)]
        return buf[off:off + int(np.prod(shp))].view(shp)

    def stat(self, name, role):
        off, shp = self.stat_entries[(name, role)]
        return self.stats[off:off + shp[0]]

    # -- Keras <-> storage ---------------------------------------------------------------------------------------
    def set_weights(self, weights, exclude=None):
        """{layer_name: [arrays in Keras layer.weights order]} -> buffers; unknown layers are skipped (by-name loading)."""
        torch = self.torch
        loaded = 0
        for name, arrays in weights.items():
            if name not in self.kinds or (exclude and name in exclude):
                continue
            kind = self.kinds[name]
            arrs = [torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(self.device) for a in arrays]
            if kind == "bn":
                self.view(self.w, name, "gamma").copy_(arrs[0])
                self.view(self.w, name, "beta").copy_(arrs[1])
                self.stat(name, "mean").copy_(arrs[2])
                self.stat(name, "var").copy_(arrs[3])
            else:
                k = arrs[0]
                if kind == "conv":
                    k = k.permute(3, 0, 1, 2)
                elif kind == "dense":
                    k = k.t()
                self.view(self.w, name, "kernel").copy_(k)
                self.view(self.w, name, "bias").copy_(arrs[1])
            loaded += 1
        self.wb.copy_(self.w)
        return loaded

    def get_weights(self):
        out = {}
        for name, kind, _ in self.specs:
            if kind == "bn":
                out[name] = [self.view(self.w, name, "gamma").cpu().numpy(), self.view(self.w, name, "beta").cpu().numpy(),
                             self.stat(name, "mean").cpu().numpy(), self.stat(name, "var").cpu().numpy()]
            else:
                k = self.view(self.w, name, "kernel")
                if kind == "conv":
                    k = k.permute(1, 2, 3, 0)
                elif kind == "dense":
                    k = k.t()
                out[name] = [k.contiguous().cpu().numpy(), self.view(self.w, name, "bias").cpu().numpy()]
        return out

    def segments(self, weight_decay, trainable):
        """(segment_start int64 [nseg+1], reg_coef float32 [nseg], lr_mask) for mrcnn_sgd_step: the L2 regulariser of
        MaskRCNN.compile is l2(WEIGHT_DECAY)(w) / size(w) on kernels and biases, nothing on gamma / beta
        (mrcnn/model.py:2281-2286); gradient = 2 * WEIGHT_DECAY * w / size(w).  Layers outside `trainable` get no
        regulariser (their gradients are never produced, so they do not move)."""
        starts, coefs = [], []
        for name, role, off, size in self.order:
            starts.append(off)
            reg = role in ("kernel", "bias") and trainable(name)
            coefs.append(2.0 * float(weight_decay) / float(size) if reg else 0.0)
        starts.append(self.n)
        return np.asarray(starts, dtype=np.int64), np.asarray(coefs, dtype=np.float32)


# ============================================================================================================
# the training graph
# ============================================================================================================

LAYER_REGEX = {
    "heads": r"(mrcnn\_.*)|(rpn\_.*)|(fpn\_.*)",
    "3+": r"(res3.*)|(bn3.*)|(res4.*)|(bn4.*)|(res5.*)|(bn5.*)|(mrcnn\_.*)|(rpn\_.*)|(fpn\_.*)",
    "4+": r"(res4.*)|(bn4.*)|(res5.*)|(bn5.*)|(mrcnn\_.*)|(rpn\_.*)|(fpn\_.*)",
    "5+": r"(res5.*)|(bn5.*)|(mrcnn\_.*)|(rpn\_.*)|(fpn\_.*)",
    "all": ".*",
}


def _functions(torch):
    """autograd Functions (built lazily: torch is imported on first use)."""

    class UseParam(torch.autograd.Function):
        """The bf16 operand copy of a float32 master parameter; its gradient flows back to the master in float32."""
        @staticmethod
        def forward(ctx, master, operand):
            return operand.detach()

        @staticmethod
        def backward(ctx, grad):
            return grad.float(), None

    class RoiAlign(torch.autograd.Function):
        """PyramidROIAlign on bf16 NHWC pyramid levels: forward csrc/roialign.cu, backward roialign_backward_kernel."""
        @staticmethod
        def forward(ctx, boxes, pool, image_area, p2, p3, p4, p5):
            lib = _native.lib()
            feats = [t.permute(0, 2, 3, 1) for t in (p2, p3, p4, p5)]         # NCHW channels-last -> NHWC views
            assert all(f.is_contiguous() for f in feats)
            B, N = boxes.shape[:2]
            C = feats[0].shape[-1]
            out = torch.empty((B, N, pool, pool, C), dtype=torch.bfloat16, device=boxes.device)
            levels = torch.empty((B, N), dtype=torch.int32, device=boxes.device)
            ptrs = (ctypes.c_void_p * 4)(*[f.data_ptr() for f in feats])
            hs = (ctypes.c_int * 4)(*[f.shape[1] for f in feats])
            ws = (ctypes.c_int * 4)(*[f.shape[2] for f in feats])
            st = torch.cuda.current_stream(boxes.device).cuda_stream
            with torch.cuda.device(boxes.device):
                _native.check(lib.mrcnn_pyramid_roi_align(ptrs, hs, ws, C, _native.DTYPE_BF16, _native.ptr(boxes), B, N, pool,
                                                          float(image_area), _native.ptr(out), _native.ptr(levels), st),
                              "pyramid_roi_align")
            ctx.save_for_backward(boxes, levels)
            ctx.meta = (pool, [tuple(f.shape) for f in feats])
            return out

        @staticmethod
        def backward(ctx, dout):
            lib = _native.lib()
            boxes, levels = ctx.saved_tensors
            pool, shapes = ctx.meta
            dout = dout.contiguous()
            B, N = boxes.shape[:2]
            grads = [torch.zeros(s, dtype=torch.float32, device=boxes.device) for s in shapes]
            ptrs = (ctypes.c_void_p * 4)(*[gd.data_ptr() for gd in grads])
            hs = (ctypes.c_int * 4)(*[s[1] for s in shapes])
            ws = (ctypes.c_int * 4)(*[s[2] for s in shapes])
            st = torch.cuda.current_stream(boxes.device).cuda_stream
            with torch.cuda.device(boxes.device):
                _native.check(lib.mrcnn_pyramid_roi_align_backward(ptrs, hs, ws, shapes[0][3], _native.ptr(boxes),
                                                                   _native.ptr(levels), B, N, pool, _native.ptr(dout), st),
                              "pyramid_roi_align_backward")
            outs = [gd.to(torch.bfloat16).permute(0, 3, 1, 2) for gd in grads]
            return (None, None, None) + tuple(outs)

    class ConvTC(torch.autograd.Function):
        """conv + folded BN affine + residual (+ nearest-2x upsampled residual) + ReLU as ONE tcgen05 implicit-GEMM launch
        (csrc/conv_gemm.cu, the detect path's kernel).  Backward: one pass over the incoming gradient
        (mrcnn_conv_backward_prep: ReLU mask, scale, shift gradient), the same GEMM kernel on the flipped / transposed
        weights for the data gradient, the MN-major tcgen05 GEMM (mrcnn_conv2d_wgrad_bf16) for the weight gradient.
        x, residual, result: NCHW-logical channels-last bf16 (= NHWC memory).  `L` carries the layer: w_op (bf16
        [Cout,KH,KW,Cin]), scale (or None) / shift float32 [Cout] WITHOUT autograd history, and where the parameter
        gradients go: g_kernel (float32 view of the flat gradient buffer; the weight gradient is accumulated there by the
        kernel itself), d_scale / d_shift (views of the graph's affine-gradient buffers), notify(key) for the reducer.
        With L.g_kernel None the weight gradient is returned to autograd for `w_master` instead (layers used several
        times per step)."""
        @staticmethod
        def forward(ctx, x, residual, w_master, L):
            lib = _native.lib()
            xn = x.permute(0, 2, 3, 1)
            assert xn.is_contiguous() and xn.dtype == torch.bfloat16
            N, H, W, Cin = xn.shape
            Cout, KH, KW, _ = L.w_op.shape
            stride, pad = L.stride, L.pad
            OH, OW = (H + 2 * pad - KH) // stride + 1, (W + 2 * pad - KW) // stride + 1
            out = torch.empty((N, OH, OW, Cout), dtype=torch.bfloat16, device=x.device)
            sc = L.scale if L.scale is not None else L.const(Cout, 1.0)
            rn = None
            if residual is not None:
                rn = residual.permute(0, 2, 3, 1)
                assert rn.is_contiguous()
            desc = _native.ConvDesc(n=N, h=H, w=W, cin=Cin, kh=KH, kw=KW, stride=stride, pad=pad, cout=Cout, relu=int(L.relu),
                                    residual_upsample2=int(L.res_up2), out_dtype=_native.DTYPE_BF16, out_mode=0, out_ld=0)
            st = torch.cuda.current_stream(x.device).cuda_stream
            with torch.cuda.device(x.device):
                _native.check(lib.mrcnn_conv2d_bf16(ctypes.byref(desc), _native.ptr(xn), _native.ptr(L.w_op), _native.ptr(sc),
                                                    _native.ptr(L.shift), _native.ptr(rn), _native.ptr(out), st), "conv2d")
            ctx.save_for_backward(xn, out if L.relu else None)
            ctx.L = L
            ctx.has_res = residual is not None
            ctx.in_shape = (N, H, W, Cin)
            return out.permute(0, 3, 1, 2)

        @staticmethod
        def backward(ctx, dout):
            lib = _native.lib()
            xn, out = ctx.saved_tensors
            L = ctx.L
            N, H, W, Cin = ctx.in_shape
            Cout, KH, KW, _ = L.w_op.shape
            stride, pad = L.stride, L.pad
            has_scale = L.scale is not None
            dev = xn.device
            dn = dout.permute(0, 2, 3, 1)
            if not dn.is_contiguous():
                dn = dn.contiguous()
            OH, OW = dn.shape[1], dn.shape[2]
            rows = N * OH * OW
            st = torch.cuda.current_stream(dev).cuda_stream
            need_dz = ctx.has_res or not has_scale
            plain = out is None and not has_scale            # no ReLU, no scale: dz is dout itself
            dz = dn if plain else (torch.empty_like(dn) if need_dz else None)
            dzs = torch.empty_like(dn) if has_scale else None
            d_shift = L.d_shift if L.d_shift is not None else torch.zeros(Cout, dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                _native.check(lib.mrcnn_conv_backward_prep(_native.ptr(dn), _native.ptr(out), _native.ptr(L.scale),
                                                           None if plain else _native.ptr(dz), _native.ptr(dzs), _native.ptr(d_shift),
                                                           rows, Cout, st), "conv_backward_prep")
            op = dzs if has_scale else dz
            d_res = None
            if ctx.has_res:
                d_res = dz.permute(0, 3, 1, 2)
                if L.res_up2:
                    d_res = (torch.nn.functional.avg_pool2d(d_res.float(), 2) * 4.0).to(torch.bfloat16)
            # weight gradient, accumulated by the kernel into the flat gradient buffer (or a scratch buffer for autograd)
            xs = xn if stride == 1 else xn[:, ::stride, ::stride, :].contiguous()
            G = L.g_kernel if L.g_kernel is not None else torch.zeros((Cout, KH, KW, Cin), dtype=torch.float32, device=dev)
            desc = _native.ConvDesc(n=N, h=OH, w=OW, cin=Cin, kh=KH, kw=KW, stride=1, pad=pad, cout=Cout, relu=0,
                                    residual_upsample2=0, out_dtype=_native.DTYPE_F32, out_mode=0, out_ld=0)
            with torch.cuda.device(dev):
                # d scale[c] = sum dz * (conv output before the affine) = <W[c], dW[c]> / scale[c] (dW already carries the
                # scale): the GEMM's epilogue leaves <W[c], dW[c]> in d_scale, finish_backward divides all layers at once
                fold = has_scale and L.d_scale is not None
                _native.check(lib.mrcnn_conv2d_wgrad_bf16(ctypes.byref(desc), _native.ptr(xs), _native.ptr(op), _native.ptr(G),
                                                          _native.ptr(L.w_op) if fold else None,
                                                          _native.ptr(L.d_scale) if fold else None, st), "conv2d_wgrad")
            if L.notify is not None:
                L.notify()
            # data gradient: correlation of (dz * scale) with the flipped, transposed filter — the forward kernel again
            d_x = None
            if ctx.needs_input_grad[0]:
                # the same implicit GEMM with the B operand read MN-major from the forward weights themselves
                # (mrcnn_conv2d_dgrad_bf16): no flipped / transposed copy of the weights
                dsub = torch.empty((N, OH, OW, Cin), dtype=torch.bfloat16, device=dev)
                desc = _native.ConvDesc(n=N, h=OH, w=OW, cin=Cin, kh=KH, kw=KW, stride=1, pad=pad, cout=Cout, relu=0,
                                        residual_upsample2=0, out_dtype=_native.DTYPE_BF16, out_mode=0, out_ld=0)
                with torch.cuda.device(dev):
                    _native.check(lib.mrcnn_conv2d_dgrad_bf16(ctypes.byref(desc), _native.ptr(op), _native.ptr(L.w_op),
                                                              _native.ptr(L.const(Cin, 1.0)), _native.ptr(L.const(Cin, 0.0)),
                                                              _native.ptr(dsub), st), "conv2d_dgrad")
                if stride == 1:
                    d_x = dsub.permute(0, 3, 1, 2)
                else:
                    full = torch.zeros((N, H, W, Cin), dtype=torch.bfloat16, device=dev)
                    full[:, ::stride, ::stride, :] = dsub
                    d_x = full.permute(0, 3, 1, 2)
            d_w = G if L.g_kernel is None else None
            return d_x, d_res, d_w, None

    return UseParam, RoiAlign, ConvTC


class _Layer(object):
    """what ConvTC needs to know about one layer (see ConvTC)"""
    __slots__ = ("w_op", "scale", "shift", "stride", "pad", "relu", "res_up2", "g_kernel", "d_scale", "d_shift", "notify", "const")


class TrainGraph(object):
    """The mode='training' graph (mrcnn/model.py:1935-2132) on one GPU: forward, the five losses, backward into the flat
    gradient buffer.  `layers` selects the trainable layers as MaskRCNN.train does (regex on layer names)."""

    def __init__(self, config, device=None, layers="all", seed=0):
        torch = utils._torch()
        if not torch.cuda.is_available():
            raise _native.NativeError("mrcnn (B200 build): training needs a CUDA device; there is no CPU path")
        self.torch = torch
        self.F = torch.nn.functional
        self.config = config
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        self.lib = _native.lib()
        h, w = config.IMAGE_SHAPE[:2]
        if h != w or h % 64:
            raise Exception("Image size must be dividable by 2 at least 6 times "
                            "to avoid fractions when downscaling and upscaling."
                            "For example, use 256, 320, 384, 448, 512, ... etc. ")
        if config.BACKBONE != "resnet101":
            raise NotImplementedError("mrcnn (B200 build): BACKBONE must be 'resnet101'")
        if not config.USE_RPN_ROIS:
            raise NotImplementedError("mrcnn (B200 build): USE_RPN_ROIS=False (externally supplied ROIs)")
        if config.TRAIN_BN:
            raise NotImplementedError("mrcnn (B200 build): TRAIN_BN=True (batch statistics); the reference trains with frozen BN")
        if config.MASK_LOSS_FUNCTION != "binary_crossentropy":
            raise NotImplementedError("mrcnn (B200 build): MASK_LOSS_FUNCTION must be 'binary_crossentropy'")
        if int(config.POST_NMS_ROIS_TRAINING) > 2048 or int(config.MAX_GT_INSTANCES) > 512:
            raise NotImplementedError("mrcnn (B200 build): POST_NMS_ROIS_TRAINING <= 2048 and MAX_GT_INSTANCES <= 512")
        self.params = ParamStore(config, torch, self.device)
        self.UseParam, self.RoiAlign, self.ConvTC = _functions(torch)
        import os
        self.backend = os.environ.get("MRCNN_B200_TRAIN_BACKEND", "tcgen05")       # "torch": library convolutions everywhere
        self.step_seed = int(seed)
        self.set_trainable(layers)
        a = utils.generate_pyramid_anchors(config.RPN_ANCHOR_SCALES, config.RPN_ANCHOR_RATIOS,
                                           utils.compute_backbone_shapes(config, config.IMAGE_SHAPE),
                                           config.BACKBONE_STRIDES, config.RPN_ANCHOR_STRIDE)
        self.anchors_px = a
        self.anchors = torch.from_numpy(np.ascontiguousarray(utils.norm_boxes(a, config.IMAGE_SHAPE[:2]), dtype=np.float32)).to(self.device)
        self._consts = {}
        self.reducer = None                  # set by GradReducer: told when a gradient was deposited outside autograd
        nb = self.params.bn_elements
        self._S = torch.zeros(nb, dtype=torch.float32, device=self.device)      # folded affine of every BN layer (per step)
        self._T = torch.zeros(nb, dtype=torch.float32, device=self.device)
        self._dS = torch.zeros(nb, dtype=torch.float32, device=self.device)     # and its gradient, filled by ConvTC.backward
        self._dT = torch.zeros(nb, dtype=torch.float32, device=self.device)
        self._box_shift = torch.tensor([0., 0., 1., 1.], device=self.device)
        self._box_scale = torch.full((4,), float(config.IMAGE_SHAPE[0]) - 1.0, device=self.device)
        self.seed_device = torch.zeros(1, dtype=torch.int64, device=self.device)     # advanced on the device every step
        self.taps = {}

    # -- trainable layers --------------------------------------------------------------------------------------
    def set_trainable(self, layers):
        rx = LAYER_REGEX.get(layers, layers)
        self._trainable_rx = re.compile(rx)
        self.trainable = lambda name: bool(self._trainable_rx.fullmatch(name))
        p = self.params
        self.masters = {}
        for name, role, off, size in p.order:
            t = p.view(p.w, name, role).detach()
            if self.trainable(name):
                t.requires_grad_(True)
                t.grad = p.view(p.g, name, role)
            self.masters[(name, role)] = t
        mask = self.torch.zeros(p.bn_elements, dtype=self.torch.float32, device=self.device)
        cmask = self.torch.zeros(p.bn_elements, dtype=self.torch.float32, device=self.device)
        for conv, bn in p.bn_of.items():
            r, c = p.bn_rel[bn]
            mask[r:r + c] = 1.0 if self.trainable(bn) else 0.0
            cmask[r:r + c] = 1.0 if self.trainable(conv) else 0.0
        self._bn_trainable, self._bnconv_trainable = mask, cmask

    def const(self, n, value):
        """cached float32 vector of n copies of value (scale 1 / shift 0 of the plain GEMM launches)"""
        key = (int(n), float(value))
        if key not in self._consts:
            self._consts[key] = self.torch.full((int(n),), float(value), dtype=self.torch.float32, device=self.device)
        return self._consts[key]

    def begin_step(self):
        """Folded affine of every BatchNorm layer in five vector operations (no autograd: ConvTC deposits its gradient in
        _dS / _dT and finish_backward turns those into the gamma / beta / bias gradients)."""
        p, torch = self.params, self.torch
        with torch.no_grad():
            torch.rsqrt(p.bn_group(None, "var") + BN_EPS, out=self._S)
            self._S.mul_(p.bn_group(p.w, "gamma"))
            torch.sub(p.bn_group(p.w, "bias_bn"), p.bn_group(None, "mean"), out=self._T)
            self._T.mul_(self._S).add_(p.bn_group(p.w, "beta"))
            self._dS.zero_()
            self._dT.zero_()

    def finish_backward(self):
        """After loss.backward(): gradients of gamma, beta and the BN'd convolutions' biases from the affine gradients the
        fused layers left in _dS / _dT:  scale = gamma*inv, shift = (bias - mean)*scale + beta, inv = rsqrt(var + eps)
          d gamma = (dS + dT*(bias - mean)) * inv ;  d beta = dT ;  d bias = dT * scale."""
        p, torch = self.params, self.torch
        with torch.no_grad():
            inv = torch.rsqrt(p.bn_group(None, "var") + BN_EPS)
            centred = p.bn_group(p.w, "bias_bn") - p.bn_group(None, "mean")
            # _dS holds <W, dW> per channel (weight-gradient epilogue); dW carries the scale once too often
            self._dS.copy_(torch.where(self._S != 0, self._dS / self._S, torch.zeros_like(self._dS)))
            p.bn_group(p.g, "gamma").add_((self._dS + self._dT * centred) * inv * self._bn_trainable)
            p.bn_group(p.g, "beta").add_(self._dT * self._bn_trainable)
            p.bn_group(p.g, "bias_bn").add_(self._dT * self._S * self._bnconv_trainable)

    def trainable_parameters(self):
        return [(k, t) for k, t in self.masters.items() if t.requires_grad]

    # -- parameter access ----------------------------------------------------------------------------------------
    def _operand(self, name, role="kernel"):
        m = self.masters[(name, role)]
        op = self.params.view(self.params.wb, name, role)
        return self.UseParam.apply(m, op) if m.requires_grad else op

    def _affine(self, conv, bn):
        """(scale, shift) float32 with y = acc*scale + shift: bias and the frozen BatchNorm folded (BN eps 1e-3)."""
        bias = self.masters[(conv, "bias")]
        if bn is None:
            return None, bias
        gamma, beta = self.masters[(bn, "gamma")], self.masters[(bn, "beta")]
        s = gamma * self.torch.rsqrt(self.params.stat(bn, "var") + BN_EPS)
        return s, (bias - self.params.stat(bn, "mean")) * s + beta

    def _finish(self, acc, conv, bn, relu, residual=None):
        s, t = self._affine(conv, bn)
        y = acc.float()
        y = y * s.view(1, -1, 1, 1) + t.view(1, -1, 1, 1) if s is not None else y + t.view(1, -1, 1, 1)
        if residual is not None:
            y = y + residual.float()
        if relu:
            y = self.F.relu(y)
        return y.to(self.torch.bfloat16).contiguous(memory_format=self.torch.channels_last)

    def _layer(self, name, bn, w_shape, relu, stride, pad, res_up2, shared):
        """the _Layer record of ConvTC for layer `name` (kernel seen as w_shape = [Cout,KH,KW,Cin])"""
        p = self.params
        L = _Layer()
        L.w_op = p.view(p.wb, name, "kernel").view(w_shape)
        L.stride, L.pad, L.relu, L.res_up2, L.const = stride, pad, relu, res_up2, self.const
        train_k = self.masters[(name, "kernel")].requires_grad
        if bn is not None:
            r, c = p.bn_rel[bn]
            L.scale, L.shift = self._S[r:r + c], self._T[r:r + c]
            L.d_scale, L.d_shift = self._dS[r:r + c], self._dT[r:r + c]
        else:
            L.scale, L.shift = None, p.view(p.w, name, "bias")
            L.d_scale = None
            L.d_shift = p.view(p.g, name, "bias") if self.masters[(name, "bias")].requires_grad else None
        # single-use layers: the weight-gradient kernel accumulates straight into the flat gradient buffer
        L.g_kernel = p.view(p.g, name, "kernel").view(w_shape) if (train_k and not shared) else None
        key = (name, "kernel")
        L.notify = (lambda: self.reducer.mark_ready(key)) if (self.reducer is not None and L.g_kernel is not None) else None
        if not train_k and not shared:            # frozen layer: its gradient goes to a scratch buffer nobody reads
            L.g_kernel = self.torch.zeros(w_shape, dtype=self.torch.float32, device=self.device)
        return L

    def conv(self, x, name, bn=None, relu=False, stride=1, pad=0, residual=None, residual_up2=False, shared=False):
        """One convolution layer with its folded BatchNorm, optional residual (nearest-2x upsampled for the FPN laterals)
        and ReLU.  tcgen05 backend: one fused launch (ConvTC) for 1x1 / 3x3 layers whose channel counts fit the
        implicit GEMM (Cin and Cout multiples of 64); everything else — the 7x7 stem with 3 input channels, the 6 / 12
        channel RPN outputs — goes through torch.  shared: the layer is applied several times per step (the RPN on five
        pyramid levels), its weight gradient is summed by autograd."""
        m = self.masters[(name, "kernel")]
        cout, kh, kw, cin = m.shape
        if (self.backend == "tcgen05" and cin % 64 == 0 and cout % 64 == 0 and (kh, kw, pad) in ((1, 1, 0), (3, 3, 1))
                and (stride == 1 or kh == 1)):
            L = self._layer(name, bn, (cout, kh, kw, cin), relu, stride, pad, residual_up2, shared)
            return self.ConvTC.apply(x, residual, m if (shared and m.requires_grad) else None, L)
        if residual is not None and residual_up2:
            residual = self.F.interpolate(residual, scale_factor=2, mode="nearest")
        w = self._operand(name).permute(0, 3, 1, 2)           # [Cout,KH,KW,Cin] storage seen as channels-last OIHW
        return self._finish(self.F.conv2d(x, w, None, stride=stride, padding=pad), name, bn, relu, residual)

    def _block(self, x, stage, blk, first, stride):
        base, bnb = "res%d%s_branch" % (stage, blk), "bn%d%s_branch" % (stage, blk)
        y = self.conv(x, base + "2a", bnb + "2a", relu=True, stride=stride)
        y = self.conv(y, base + "2b", bnb + "2b", relu=True, pad=1)
        sc = self.conv(x, base + "1", bnb + "1", stride=stride) if first else x
        return self.conv(y, base + "2c", bnb + "2c", relu=True, residual=sc)

    def backbone_fpn(self, images):
        """images [B,S,S,3] float32 (molded) -> P2..P6, NCHW-logical channels-last bf16 (mrcnn/model.py:175-210, 2003-2026)."""
        torch, F = self.torch, self.F
        x = images.to(torch.bfloat16).permute(0, 3, 1, 2)
        x = self.conv(x, "conv1", "bn_conv1", relu=True, stride=2, pad=3)
        x = F.max_pool2d(F.pad(x, (0, 1, 0, 1), value=float("-inf")), 3, 2)      # MaxPooling2D(3, 2, 'same')
        feats = {}
        for stage, nblocks in ((2, 3), (3, 4), (4, 23), (5, 3)):
            for i in range(nblocks):
                x = self._block(x, stage, chr(97 + i), i == 0, 2 if (i == 0 and stage > 2) else 1)
            feats["C%d" % stage] = x
        p5 = self.conv(feats["C5"], "fpn_c5p5")
        p4 = self.conv(feats["C4"], "fpn_c4p4", residual=p5, residual_up2=True)
        p3 = self.conv(feats["C3"], "fpn_c3p3", residual=p4, residual_up2=True)
        p2 = self.conv(feats["C2"], "fpn_c2p2", residual=p3, residual_up2=True)
        P = {"P2": self.conv(p2, "fpn_p2", pad=1), "P3": self.conv(p3, "fpn_p3", pad=1),
             "P4": self.conv(p4, "fpn_p4", pad=1), "P5": self.conv(p5, "fpn_p5", pad=1)}
        P["P6"] = F.max_pool2d(P["P5"], 1, 2)
        return P

    def rpn(self, P):
        """-> rpn_class_logits [B,A,2], rpn_class [B,A,2], rpn_bbox [B,A,4] float32 (mrcnn/model.py:916-957, 2040-2055)."""
        torch = self.torch
        logits, boxes = [], []
        for lvl in ("P2", "P3", "P4", "P5", "P6"):
            s = self.conv(P[lvl], "rpn_conv_shared", relu=True, pad=1, shared=True)
            c = self.conv(s, "rpn_class_raw").permute(0, 2, 3, 1)
            d = self.conv(s, "rpn_bbox_pred").permute(0, 2, 3, 1)
            logits.append(c.reshape(c.shape[0], -1, 2).float())
            boxes.append(d.reshape(d.shape[0], -1, 4).float())
        rpn_class_logits = torch.cat(logits, 1)
        return rpn_class_logits, torch.softmax(rpn_class_logits, -1), torch.cat(boxes, 1)

    def proposals(self, rpn_class, rpn_bbox):
        """ProposalLayer with POST_NMS_ROIS_TRAINING rows (mrcnn/model.py:2060-2066, 329-406); no gradient."""
        torch, cfg = self.torch, self.config
        B, A = rpn_class.shape[:2]
        R = int(cfg.POST_NMS_ROIS_TRAINING)
        K = min(int(cfg.PRE_NMS_LIMIT), A)
        rois = torch.empty((B, R, 4), dtype=torch.float32, device=self.device)
        topk = torch.empty((B, K), dtype=torch.int32, device=self.device)
        sd = _native.float_array([float(np.float32(v)) for v in cfg.RPN_BBOX_STD_DEV])
        rc, rb = rpn_class.detach().contiguous(), rpn_bbox.detach().contiguous()
        with torch.cuda.device(self.device):
            _native.check(self.lib.mrcnn_proposal_layer(_native.ptr(rc), _native.ptr(rb), _native.ptr(self.anchors), 0, B, A,
                                                        int(cfg.PRE_NMS_LIMIT), R, float(cfg.RPN_NMS_THRESHOLD), sd,
                                                        _native.ptr(rois), _native.ptr(topk), None, None, None, 0,
                                                        torch.cuda.current_stream(self.device).cuda_stream), "proposal_layer")
        return rois

    def detection_targets(self, rpn_rois, gt_class_ids, gt_boxes_norm, gt_masks, seed, seed_device=None):
        """DetectionTargetLayer (mrcnn/model.py:570-763) -> rois [B,T,4], target_class_ids [B,T] int32, target_bbox
        [B,T,4], target_mask [B,T,mh,mw] float32, counts [B,2]."""
        torch, cfg = self.torch, self.config
        B, N = rpn_rois.shape[:2]
        T = int(cfg.TRAIN_ROIS_PER_IMAGE)
        mh, mw = int(cfg.MASK_SHAPE[0]), int(cfg.MASK_SHAPE[1])
        G = gt_class_ids.shape[1]
        rois = torch.empty((B, T, 4), dtype=torch.float32, device=self.device)
        tcls = torch.empty((B, T), dtype=torch.int32, device=self.device)
        tbox = torch.empty((B, T, 4), dtype=torch.float32, device=self.device)
        tmask = torch.empty((B, T, mh, mw), dtype=torch.float32, device=self.device)
        counts = torch.empty((B, 2), dtype=torch.int32, device=self.device)
        sd = _native.float_array([float(np.float32(v)) for v in cfg.BBOX_STD_DEV])
        gm = gt_masks.contiguous()
        with torch.cuda.device(self.device):
            _native.check(self.lib.mrcnn_detection_targets(
                _native.ptr(rpn_rois), _native.ptr(gt_class_ids), _native.ptr(gt_boxes_norm), _native.ptr(gm), B, N, G,
                gm.shape[1], gm.shape[2], 1 if cfg.USE_MINI_MASK else 0, T, float(cfg.ROI_POSITIVE_RATIO), sd, mh, mw,
                int(seed), _native.ptr(seed_device), _native.ptr(rois), _native.ptr(tcls), _native.ptr(tbox), _native.ptr(tmask),
                _native.ptr(counts),
                torch.cuda.current_stream(self.device).cuda_stream), "detection_targets")
        return rois, tcls, tbox, tmask, counts

    def class_head(self, rois, P):
        """fpn_classifier_graph (mrcnn/model.py:986-1039) -> logits [B,T,NC], probs, bbox deltas [B,T,NC,4] float32."""
        torch, cfg = self.torch, self.config
        B, T = rois.shape[:2]
        area = float(cfg.IMAGE_SHAPE[0] * cfg.IMAGE_SHAPE[1])
        x = self.RoiAlign.apply(rois, int(cfg.POOL_SIZE), area, P["P2"], P["P3"], P["P4"], P["P5"])
        x = x.view(B * T, cfg.POOL_SIZE, cfg.POOL_SIZE, -1).permute(0, 3, 1, 2)
        x = self._fc1(x)
        x = self.conv(x, "mrcnn_class_conv2", "mrcnn_class_bn2", relu=True)
        shared = x.reshape(B * T, -1)
        logits = (shared @ self._operand("mrcnn_class_logits").t()).float() + self.masters[("mrcnn_class_logits", "bias")]
        bbox = (shared @ self._operand("mrcnn_bbox_fc").t()).float() + self.masters[("mrcnn_bbox_fc", "bias")]
        nc = int(cfg.NUM_CLASSES)
        logits = logits.view(B, T, nc)
        return logits, torch.softmax(logits, -1), bbox.view(B, T, nc, 4)

    def _fc1(self, x):
        """mrcnn_class_conv1: a POOL x POOL VALID convolution over a POOL x POOL map = one dense layer over K = POOL*POOL*C;
        the pooled tensor [n, P, P, C] and the kernel [Cout, P, P, C] are already laid out as that GEMM's operands."""
        name, bn = "mrcnn_class_conv1", "mrcnn_class_bn1"
        cout, kh, kw, cin = self.masters[(name, "kernel")].shape
        k = kh * kw * cin
        if self.backend == "tcgen05" and k % 64 == 0 and cout % 64 == 0:
            xf = x.permute(0, 2, 3, 1).reshape(x.shape[0], 1, 1, k).permute(0, 3, 1, 2)
            L = self._layer(name, bn, (cout, 1, 1, k), True, 1, 0, False, False)
            return self.ConvTC.apply(xf, None, None, L)
        return self.conv(x, name, bn, relu=True)

    def mask_head(self, rois, P):
        """build_fpn_mask_graph (mrcnn/model.py:1042-1091) -> masks [B,T,2*MASK_POOL,2*MASK_POOL,NC] float32 in (0,1)."""
        torch, cfg, F = self.torch, self.config, self.F
        B, T = rois.shape[:2]
        mp = int(cfg.MASK_POOL_SIZE)
        area = float(cfg.IMAGE_SHAPE[0] * cfg.IMAGE_SHAPE[1])
        x = self.RoiAlign.apply(rois, mp, area, P["P2"], P["P3"], P["P4"], P["P5"])
        x = x.view(B * T, mp, mp, -1).permute(0, 3, 1, 2)
        for i in range(1, 5):
            x = self.conv(x, "mrcnn_mask_conv%d" % i, "mrcnn_mask_bn%d" % i, relu=True, pad=1)
        # transposed convolution + ReLU and the NUM_CLASSES-channel 1x1 through torch (library), bias added in the call
        wd = self._operand("mrcnn_mask_deconv").permute(3, 2, 0, 1)          # [KH,KW,Cout,Cin] -> [Cin,Cout,KH,KW]
        x = F.relu(F.conv_transpose2d(x, wd, self.masters[("mrcnn_mask_deconv", "bias")].to(torch.bfloat16), stride=2))
        x = F.conv2d(x, self._operand("mrcnn_mask").permute(0, 3, 1, 2), self.masters[("mrcnn_mask", "bias")].to(torch.bfloat16))
        x = torch.sigmoid(x.float())
        return x.permute(0, 2, 3, 1).reshape(B, T, 2 * mp, 2 * mp, int(cfg.NUM_CLASSES))

    # -- losses (mrcnn/model.py:1098-1270), float32 ------------------------------------------------------------------
    def losses(self, rpn_match, rpn_bbox_t, rpn_class_logits, rpn_bbox, target_class_ids, target_bbox, target_mask,
               mrcnn_class_logits, mrcnn_bbox, mrcnn_mask, active_class_ids):
        """The five loss graphs as masked sums: a mean over the gathered rows of the reference is sum(mask * loss) /
        count, and "no rows -> 0" falls out of count clamped to 1 — no host decision anywhere, so the step can be
        captured in a CUDA graph."""
        torch, F = self.torch, self.F

        def smooth_l1(t, p):
            d = (t - p).abs()
            return torch.where(d < 1.0, 0.5 * d * d, d - 0.5)

        out = {}
        match = rpn_match.view(rpn_match.shape[0], -1)
        used = (match != 0).float()
        ce = F.cross_entropy(rpn_class_logits.reshape(-1, 2), (match == 1).long().reshape(-1), reduction="none")
        out["rpn_class_loss"] = (ce * used.reshape(-1)).sum() / used.sum().clamp(min=1.0)
        posm = match == 1
        # batch_pack_graph: the k-th positive anchor of image b (anchor order) goes with target row k of image b
        row = (posm.long().cumsum(1) - 1).clamp(0, rpn_bbox_t.shape[1] - 1)
        tgt = rpn_bbox_t.float().gather(1, row.unsqueeze(-1).expand(-1, -1, 4))
        pf = posm.float()
        out["rpn_bbox_loss"] = (smooth_l1(tgt, rpn_bbox) * pf.unsqueeze(-1)).sum() / (4.0 * pf.sum()).clamp(min=1.0)
        tci = target_class_ids.long().reshape(-1)
        nc = mrcnn_class_logits.shape[-1]
        ce = F.cross_entropy(mrcnn_class_logits.reshape(-1, nc), tci, reduction="none")
        pred_active = active_class_ids[0].float()[mrcnn_class_logits.argmax(-1).reshape(-1)]
        out["mrcnn_class_loss"] = (ce * pred_active).sum() / pred_active.sum()
        pos = (tci > 0).float()
        npos = pos.sum()
        pred = mrcnn_bbox.reshape(-1, nc, 4).gather(1, tci.view(-1, 1, 1).expand(-1, 1, 4)).squeeze(1)
        out["mrcnn_bbox_loss"] = (smooth_l1(target_bbox.reshape(-1, 4), pred) * pos.unsqueeze(-1)).sum() / (4.0 * npos).clamp(min=1.0)
        mm = mrcnn_mask.reshape((-1,) + tuple(mrcnn_mask.shape[2:]))
        mh, mw = mm.shape[1], mm.shape[2]
        y_pred = mm.gather(3, tci.view(-1, 1, 1, 1).expand(-1, mh, mw, 1)).squeeze(3)
        y_true = target_mask.reshape(-1, mh, mw)
        # K.binary_crossentropy of Keras 2.2.4 / TF backend: clip to [eps, 1-eps], back to logits, sigmoid CE
        o = y_pred.clamp(1e-7, 1.0 - 1e-7)
        z = torch.log(o / (1.0 - o))
        bce = F.binary_cross_entropy_with_logits(z, y_true, reduction="none")
        out["mrcnn_mask_loss"] = (bce * pos.view(-1, 1, 1)).sum() / (float(mh * mw) * npos).clamp(min=1.0)
        return out

    # -- one replica step -----------------------------------------------------------------------------------------------
    def forward(self, inputs, seed=None):
        """inputs: the data_generator list as device tensors.  -> (total loss tensor, {loss name: tensor})."""
        torch, cfg = self.torch, self.config
        images, image_meta, rpn_match, rpn_bbox_t, gt_class_ids, gt_boxes, gt_masks = inputs
        S = float(cfg.IMAGE_SHAPE[0])
        # norm_boxes_graph (mrcnn/model.py:3003-3017): (boxes - [0,0,1,1]) / (S-1)
        # divisor as a TENSOR: torch turns division by a Python scalar into a multiplication by its reciprocal (1 ulp off
        # the reference's tf.divide)
        gt_norm = ((gt_boxes.float() - self._box_shift) / self._box_scale).contiguous()
        self.begin_step()
        P = self.backbone_fpn(images)
        rpn_class_logits, rpn_class, rpn_bbox = self.rpn(P)
        rpn_rois = self.proposals(rpn_class, rpn_bbox)
        # shuffle seed of the DetectionTargetLayer: an explicit one (tests), else base + a device counter that the step
        # advances itself (so a replayed CUDA graph draws new samples every step)
        seed_dev = None
        if seed is None:
            seed, seed_dev = self.step_seed, self.seed_device
        rois, tcls, tbox, tmask, counts = self.detection_targets(rpn_rois, gt_class_ids.to(torch.int32).contiguous(), gt_norm,
                                                                 gt_masks.to(torch.uint8), seed, seed_dev)
        if seed_dev is not None:
            seed_dev.add_(1)
        logits, probs, bbox = self.class_head(rois, P)
        masks = self.mask_head(rois, P)
        active = image_meta[:, 12:].to(torch.int32)
        ls = self.losses(rpn_match, rpn_bbox_t, rpn_class_logits, rpn_bbox, tcls, tbox, tmask, logits, bbox, masks, active)
        total = None
        for name in LOSS_NAMES:
            if not cfg.USE_LOSSES.get(name, True):
                continue
            term = ls[name] * float(cfg.LOSS_WEIGHTS.get(name, 1.))
            total = term if total is None else total + term
        self.taps = {"rpn_rois": rpn_rois, "rois": rois, "target_class_ids": tcls, "target_bbox": tbox, "target_mask": tmask,
                     "counts": counts, "rpn_class_logits": rpn_class_logits, "rpn_bbox": rpn_bbox, "mrcnn_class_logits": logits,
                     "mrcnn_bbox": bbox, "mrcnn_mask": masks}
        return total, ls

    def to_device(self, inputs):
        """data_generator arrays -> device tensors (images through pinned memory)."""
        torch = self.torch
        out = []
        for a in inputs:
            t = torch.from_numpy(np.ascontiguousarray(a))
            if t.dtype == torch.float64:
                t = t.float()
            out.append(t.to(self.device, non_blocking=True))
        return out


# ============================================================================================================
# gradient all-reduce + optimiser
# ============================================================================================================

class GradReducer(object):
    """Replaces ParallelModel (mrcnn/parallel_model.py:54-104: towers on every GPU of one process, outputs merged on the
    CPU) by one process per GPU and a bucketed all-reduce of the flat gradient buffer.  Buckets of ~bucket_bytes are cut
    from the end of the buffer (the layers that finish their backward first); a bucket is launched on the process group
    as soon as every parameter in it has its gradient (post-accumulate hooks), so the transfers run under the rest of
    the backward pass.  The mean over replicas is taken by the optimiser kernel (grad_scale = 1/world)."""

    def __init__(self, graph, bucket_bytes=25 * 2 ** 20, group=None):
        import torch.distributed as dist
        self.dist, self.group, self.graph = dist, group, graph
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        p = graph.params
        bucket_elems = max(1, int(bucket_bytes) // 4)
        # bucket boundaries on tensor starts, walking the layout backwards
        bounds = [p.n]
        for name, role, off, size in reversed(p.order):
            if bounds[-1] - off >= bucket_elems:
                bounds.append(off)
        if bounds[-1] != 0:
            bounds.append(0)
        self.buckets = [(bounds[i + 1], bounds[i]) for i in range(len(bounds) - 1)]        # [(start, end)], first = last layers
        self._bucket_of = {}
        self._need = [0] * len(self.buckets)
        for (name, role), t in graph.masters.items():
            if not t.requires_grad:
                continue
            off = p.entries[(name, role)][0]
            bi = next(i for i, (s, e) in enumerate(self.buckets) if s <= off < e)
            self._bucket_of[(name, role)] = bi
            self._need[bi] += 1
            t.register_post_accumulate_grad_hook(self._make_hook(bi))
        self._left = list(self._need)
        self._works = []
        self.enabled = True
        graph.reducer = self

    def _make_hook(self, bi):
        def hook(_param):
            self._ready(bi)
        return hook

    def _ready(self, bi):
        if not self.enabled:
            return
        self._left[bi] -= 1
        if self._left[bi] == 0:
            self._launch(bi)

    def mark_ready(self, key):
        """a gradient that did not come through autograd (ConvTC lets the weight-gradient kernel accumulate straight into
        the flat buffer) is complete"""
        bi = self._bucket_of.get(key)
        if bi is not None:
            self._ready(bi)

    def _launch(self, bi):
        if self.world > 1:
            s, e = self.buckets[bi]
            self._works.append(self.dist.all_reduce(self.graph.params.g[s:e], group=self.group, async_op=True))

    def finish(self):
        """After backward: launch buckets whose parameters produced no gradient this step, wait for everything."""
        if not self.enabled:
            return
        for bi, left in enumerate(self._left):
            if left > 0:
                self._launch(bi)
        for wk in self._works:
            wk.wait()
        self._works = []
        self._left = list(self._need)

    def allreduce_only(self):
        """The same buckets back to back with nothing to hide behind (calibration of the all-reduce time)."""
        for bi in range(len(self.buckets)):
            self._launch(bi)
        for wk in self._works:
            wk.wait()
        self._works = []


class SGD(object):
    """keras.optimizers.SGD(lr, momentum, clipnorm) + the L2 regulariser of MaskRCNN.compile (mrcnn/model.py:2259-2297) as
    two kernel launches over the flat buffers (csrc/train_ops.cu)."""

    def __init__(self, graph, learning_rate, momentum, clipnorm, weight_decay, world=1):
        torch = graph.torch
        self.graph, self.lr, self.momentum, self.clipnorm, self.world = graph, float(learning_rate), float(momentum), float(clipnorm or 0.0), world
        starts, coefs = graph.params.segments(weight_decay, graph.trainable)
        self.seg_start = torch.from_numpy(starts).to(graph.device)
        self.seg_coef = torch.from_numpy(coefs).to(graph.device)
        self.nseg = len(coefs)
        self.sumsq = torch.zeros(1, dtype=torch.float64, device=graph.device)

    def step(self):
        g, torch = self.graph, self.graph.torch
        p = g.params
        with torch.cuda.device(g.device):
            _native.check(g.lib.mrcnn_sgd_step(_native.ptr(p.g), _native.ptr(p.w), _native.ptr(p.v), _native.ptr(p.wb), p.n,
                                               _native.ptr(self.seg_start), _native.ptr(self.seg_coef), self.nseg,
                                               1.0 / self.world, self.clipnorm, self.lr, self.momentum, _native.ptr(self.sumsq),
                                               torch.cuda.current_stream(g.device).cuda_stream), "sgd_step")

    def grad_norm(self):
        """global norm of the last step's (regularised, averaged) gradient — before clipping"""
        return float(self.sumsq.sqrt().item())


class Trainer(object):
    """One replica's training loop body: forward, backward with overlapped all-reduce, optimiser."""

    def __init__(self, graph, learning_rate=None, momentum=None, bucket_bytes=25 * 2 ** 20, group=None):
        cfg = graph.config
        self.graph = graph
        self.reducer = GradReducer(graph, bucket_bytes=bucket_bytes, group=group)
        self.opt = SGD(graph, cfg.LEARNING_RATE if learning_rate is None else learning_rate,
                       cfg.LEARNING_MOMENTUM if momentum is None else momentum, cfg.GRADIENT_CLIP_NORM, cfg.WEIGHT_DECAY,
                       world=self.reducer.world)
        self.timing = False            # record CUDA events around backward / all-reduce wait / optimiser
        self._events = []

    def broadcast_parameters(self):
        if self.reducer.world > 1:
            p = self.graph.params
            for buf in (p.w, p.v, p.stats):
                self.reducer.dist.broadcast(buf, src=0, group=self.reducer.group)
            p.wb.copy_(p.w)

    def capture(self, example_inputs):
        """Records the whole step (forward, backward with the bucket all-reduces launched from the hooks, optimiser) into a
        CUDA graph over static input tensors; train_step then copies a batch in and replays ~1 000 launches with one
        call.  Parameters and the shuffle counter are restored after the warm-up runs a capture needs.  Returns False (and
        stays eager) if the capture fails, e.g. a process-group backend that cannot be captured."""
        g = self.graph
        torch, p = g.torch, g.params
        self._static = [t.clone() for t in example_inputs]
        keep = [p.w.clone(), p.v.clone(), p.wb.clone(), g.seed_device.clone()]
        cur = torch.cuda.current_stream(g.device)
        side = torch.cuda.Stream(device=g.device)
        side.wait_stream(cur)
        ok = True
        try:
            with torch.cuda.stream(side):
                for _ in range(3):
                    self._eager_step(self._static, None)
            cur.wait_stream(side)
            torch.cuda.synchronize(g.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                self._static_losses = self._eager_step(self._static, None)
            self._graph = graph
        except Exception as e:            # noqa: BLE001 — stay eager, say why
            logging.warning("CUDA-graph capture of the training step failed (%s); running eagerly", e)
            self._graph, ok = None, False
            torch.cuda.synchronize(g.device)
        for dst, src in zip((p.w, p.v, p.wb, g.seed_device), keep):
            dst.copy_(src)
        return ok

    def train_step(self, inputs, seed=None):
        """-> {loss name: float tensor (this replica)}; gradients averaged over the replicas, parameters updated."""
        if getattr(self, "_graph", None) is not None and seed is None:
            for dst, src in zip(self._static, inputs):
                dst.copy_(src, non_blocking=True)
            self._graph.replay()
            return self._static_losses
        return self._eager_step(inputs, seed)

    def _eager_step(self, inputs, seed):
        g = self.graph
        torch = g.torch
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)] if self.timing else None
        if ev:
            ev[0].record()
        g.params.g.zero_()
        total, ls = g.forward(inputs, seed=seed)
        if ev:
            ev[1].record()
        total.backward()
        g.finish_backward()
        if ev:
            ev[2].record()
        self.reducer.finish()
        if ev:
            ev[3].record()
        self.opt.step()
        if ev:
            ev[4].record()
            self._events.append(ev)
        ls = dict(ls)
        ls["loss"] = total.detach()
        return ls

    def phase_ms(self):
        """mean ms of (forward, backward incl. the overlapped all-reduce launches, exposed all-reduce wait, optimiser) over
        the steps recorded with timing=True; call after torch.cuda.synchronize()."""
        if not self._events:
            return None
        names = ["forward", "backward", "allreduce_exposed", "optimizer"]
        out = {n: 0.0 for n in names}
        for ev in self._events:
            for k, n in enumerate(names):
                out[n] += ev[k].elapsed_time(ev[k + 1])
        return {n: v / len(self._events) for n, v in out.items()}
