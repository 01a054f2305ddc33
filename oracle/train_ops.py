"""ORACLE (test infrastructure — only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this).

CPU restatement of the reference's TRAINING path (BASELINE.json configs[4]):

  compute_overlaps / box_refinement   <- mrcnn/utils.py:75-97, 147-163, 275-298           (numpy; pinned by goldens)
  build_rpn_targets                   <- mrcnn/model.py:1536-1644                         (numpy; pinned by goldens)
  detection_targets                   <- mrcnn/model.py:540-683 detection_targets_graph   (TF graph: parity UNPINNED)
  losses                              <- mrcnn/model.py:1098-1270                         (TF / Keras: parity UNPINNED)
  TrainNet                            <- mrcnn/model.py:99-210, 916-1091, 2003-2132       (torch-CPU fp32 autograd)
  sgd_step                            <- keras.optimizers.SGD + clipnorm (Keras 2.2.4 optimizers.py: clip_norm on the
                                         global norm, v = m*v - lr*g, p += v) + MaskRCNN.compile's L2 regulariser
                                         (mrcnn/model.py:2259-2297)                       (third party: parity UNPINNED)

Pinned parts are checked against tests/golden/training_golden.npz (outputs of the reference's own functions).  TF pieces
are restated from the graph code; tf.random.shuffle has no reproducible order, so the shuffle is DEFINED here as
"ascending shuffle_key" — the product kernel uses the same definition.  float32 with one rounding per op; log is "double
log, one rounding to float" (the convention of oracle/graph_layers.py)."""
import numpy as np
import torch
import torch.nn.functional as F

from . import graph_layers as GL

f32 = np.float32
BN_EPS = 1e-3


# --------------------------------------------------------------------------------------------------------------
# pinned host functions
# --------------------------------------------------------------------------------------------------------------

def compute_iou(box, boxes, box_area, boxes_area):
    y1 = np.maximum(box[0], boxes[:, 0])
    y2 = np.minimum(box[2], boxes[:, 2])
    x1 = np.maximum(box[1], boxes[:, 1])
    x2 = np.minimum(box[3], boxes[:, 3])
    inter = np.maximum(x2 - x1, 0) * np.maximum(y2 - y1, 0)
    return inter / (box_area + boxes_area[:] - inter[:])


def compute_overlaps(boxes1, boxes2):
    """column i = IoU of boxes2[i] with every box of boxes1 (the reference's loop, utils.py:147-163)"""
    area1 = (boxes1[:, 2] - boxes1[:, 0]) * (boxes1[:, 3] - boxes1[:, 1])
    area2 = (boxes2[:, 2] - boxes2[:, 0]) * (boxes2[:, 3] - boxes2[:, 1])
    out = np.zeros((boxes1.shape[0], boxes2.shape[0]))
    for i in range(out.shape[1]):
        out[:, i] = compute_iou(boxes2[i], boxes1, area2[i], area1)
    return out


def box_refinement(box, gt_box):
    box, gt_box = box.astype(f32), gt_box.astype(f32)
    h, w = box[:, 2] - box[:, 0], box[:, 3] - box[:, 1]
    cy, cx = box[:, 0] + 0.5 * h, box[:, 1] + 0.5 * w
    gh, gw = gt_box[:, 2] - gt_box[:, 0], gt_box[:, 3] - gt_box[:, 1]
    gcy, gcx = gt_box[:, 0] + 0.5 * gh, gt_box[:, 1] + 0.5 * gw
    return np.stack([(gcy - cy) / h, (gcx - cx) / w, np.log(gh / h), np.log(gw / w)], axis=1)


def build_rpn_targets(image_shape, anchors, gt_class_ids, gt_boxes, config):
    """Statement-for-statement restatement (per-anchor loop kept), mrcnn/model.py:1536-1644."""
    rpn_match = np.zeros([anchors.shape[0]], dtype=np.int32)
    rpn_bbox = np.zeros((config.RPN_TRAIN_ANCHORS_PER_IMAGE, 4))
    crowd_ix = np.where(gt_class_ids < 0)[0]
    if crowd_ix.shape[0] > 0:
        non_crowd_ix = np.where(gt_class_ids > 0)[0]
        crowd_boxes = gt_boxes[crowd_ix]
        gt_class_ids = gt_class_ids[non_crowd_ix]
        gt_boxes = gt_boxes[non_crowd_ix]
        no_crowd_bool = np.amax(compute_overlaps(anchors, crowd_boxes), axis=1) < 0.001
    else:
        no_crowd_bool = np.ones([anchors.shape[0]], dtype=bool)
    overlaps = compute_overlaps(anchors, gt_boxes)
    anchor_iou_argmax = np.argmax(overlaps, axis=1)
    anchor_iou_max = overlaps[np.arange(overlaps.shape[0]), anchor_iou_argmax]
    rpn_match[(anchor_iou_max < 0.3) & no_crowd_bool] = -1
    rpn_match[np.argwhere(overlaps == np.max(overlaps, axis=0))[:, 0]] = 1
    rpn_match[anchor_iou_max >= 0.7] = 1
    ids = np.where(rpn_match == 1)[0]
    extra = len(ids) - (config.RPN_TRAIN_ANCHORS_PER_IMAGE // 2)
    if extra > 0:
        rpn_match[np.random.choice(ids, extra, replace=False)] = 0
    ids = np.where(rpn_match == -1)[0]
    extra = len(ids) - (config.RPN_TRAIN_ANCHORS_PER_IMAGE - np.sum(rpn_match == 1))
    if extra > 0:
        rpn_match[np.random.choice(ids, extra, replace=False)] = 0
    ids = np.where(rpn_match == 1)[0]
    ix = 0
    for i, a in zip(ids, anchors[ids]):
        gt = gt_boxes[anchor_iou_argmax[i]]
        gt_h, gt_w = gt[2] - gt[0], gt[3] - gt[1]
        gt_cy, gt_cx = gt[0] + 0.5 * gt_h, gt[1] + 0.5 * gt_w
        a_h, a_w = a[2] - a[0], a[3] - a[1]
        a_cy, a_cx = a[0] + 0.5 * a_h, a[1] + 0.5 * a_w
        rpn_bbox[ix] = [(gt_cy - a_cy) / a_h, (gt_cx - a_cx) / a_w, np.log(gt_h / a_h), np.log(gt_w / a_w)]
        rpn_bbox[ix] /= config.RPN_BBOX_STD_DEV
        ix += 1
    return rpn_match, rpn_bbox


# --------------------------------------------------------------------------------------------------------------
# DetectionTargetLayer
# --------------------------------------------------------------------------------------------------------------

def shuffle_key(seed, image, stream, index):
    """The sort key that defines a shuffle (same integer hash as csrc/train_ops.cu: shuffle_key), uint32 arithmetic."""
    M = 0xFFFFFFFF
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    x = ((seed & M) ^ (((seed >> 32) * 0x9E3779B9) & M)) & M
    x ^= (int(image) * 0x85EBCA6B + 0x27D4EB2F) & M
    x ^= (int(stream) * 0xC2B2AE35) & M
    x ^= (int(index) * 0x165667B1 + 0x9E3779B9) & M
    x ^= x >> 16
    x = (x * 0x85EBCA6B) & M
    x ^= x >> 13
    x = (x * 0xC2B2AE35) & M
    x ^= x >> 16
    return x


def overlaps_graph(b1, b2):
    """mrcnn/model.py:540-567, float32"""
    b1, b2 = b1.astype(f32)[:, None, :], b2.astype(f32)[None, :, :]
    y1, x1 = np.maximum(b1[..., 0], b2[..., 0]), np.maximum(b1[..., 1], b2[..., 1])
    y2, x2 = np.minimum(b1[..., 2], b2[..., 2]), np.minimum(b1[..., 3], b2[..., 3])
    inter = np.maximum(x2 - x1, f32(0)) * np.maximum(y2 - y1, f32(0))
    a1 = (b1[..., 2] - b1[..., 0]) * (b1[..., 3] - b1[..., 1])
    a2 = (b2[..., 2] - b2[..., 0]) * (b2[..., 3] - b2[..., 1])
    with np.errstate(divide="ignore", invalid="ignore"):
        return (inter / (a1 + a2 - inter)).astype(f32)


def _crop_and_resize_1ch(img, box, out_h, out_w):
    """tf.image.crop_and_resize (bilinear, extrapolation 0) of one [H,W] float32 image for one box, float32."""
    H, W = img.shape
    y1, x1, y2, x2 = [f32(v) for v in box]
    out = np.zeros((out_h, out_w), dtype=f32)
    Hm1, Wm1 = f32(H - 1), f32(W - 1)
    sy = (y2 - y1) * Hm1 / f32(out_h - 1) if out_h > 1 else f32(0)
    sx = (x2 - x1) * Wm1 / f32(out_w - 1) if out_w > 1 else f32(0)
    for iy in range(out_h):
        in_y = y1 * Hm1 + f32(iy) * sy if out_h > 1 else f32(0.5) * (y1 + y2) * Hm1
        if not (in_y >= 0 and in_y <= Hm1):
            continue
        t, b = int(np.floor(in_y)), int(np.ceil(in_y))
        ly = f32(in_y - f32(t))
        for ix in range(out_w):
            in_x = x1 * Wm1 + f32(ix) * sx if out_w > 1 else f32(0.5) * (x1 + x2) * Wm1
            if not (in_x >= 0 and in_x <= Wm1):
                continue
            l, r = int(np.floor(in_x)), int(np.ceil(in_x))
            lx = f32(in_x - f32(l))
            top = f32(img[t, l] + f32((img[t, r] - img[t, l]) * lx))
            bot = f32(img[b, l] + f32((img[b, r] - img[b, l]) * lx))
            out[iy, ix] = f32(top + f32((bot - top) * ly))
    return out


def detection_targets(proposals, gt_class_ids, gt_boxes, gt_masks, train_rois, positive_ratio, bbox_std_dev, mask_shape,
                      use_mini_mask, seed, image_index):
    """One image of detection_targets_graph (mrcnn/model.py:570-683).  proposals [N,4] / gt_boxes [G,4] normalised float32,
    gt_masks [H,W,G] bool.  -> rois [T,4], class_ids [T] int32, deltas [T,4], masks [T,mh,mw] float32, (npos, nneg)."""
    proposals, gt_boxes = np.asarray(proposals, f32), np.asarray(gt_boxes, f32)
    keep_p = np.abs(proposals).sum(axis=1) != 0                       # trim_zeros_graph
    proposals = proposals[keep_p]
    nz = np.abs(gt_boxes).sum(axis=1) != 0
    gt_boxes, gt_class_ids, gt_masks = gt_boxes[nz], np.asarray(gt_class_ids)[nz], gt_masks[:, :, nz]
    crowd = gt_boxes[gt_class_ids < 0]
    nc = gt_class_ids > 0
    gt_class_ids, gt_boxes, gt_masks = gt_class_ids[nc], gt_boxes[nc], gt_masks[:, :, nc]
    T = int(train_rois)
    mh, mw = mask_shape
    ov = overlaps_graph(proposals, gt_boxes)
    iou_max = ov.max(axis=1) if ov.shape[1] else np.full(len(proposals), -np.inf, f32)
    crowd_max = overlaps_graph(proposals, crowd).max(axis=1) if len(crowd) else np.full(len(proposals), -np.inf, f32)
    pos_ix = np.where(iou_max >= f32(0.5))[0]
    neg_ix = np.where((iou_max < f32(0.5)) & (crowd_max < f32(0.001)))[0]
    pos_ix = np.array(sorted(pos_ix, key=lambda i: (shuffle_key(seed, image_index, 0, i), i)), dtype=np.int64)
    neg_ix = np.array(sorted(neg_ix, key=lambda i: (shuffle_key(seed, image_index, 1, i), i)), dtype=np.int64)
    pos_ix = pos_ix[:int(T * positive_ratio)]
    pc = len(pos_ix)
    r = f32(1.0 / positive_ratio)
    neg_count = int(f32(r * f32(pc))) - pc
    neg_ix = neg_ix[:max(neg_count, 0)]
    rois = np.zeros((T, 4), f32)
    cls = np.zeros((T,), np.int32)
    deltas = np.zeros((T, 4), f32)
    masks = np.zeros((T, mh, mw), f32)
    if pc:
        pr = proposals[pos_ix]
        assign = ov[pos_ix].argmax(axis=1)
        g = gt_boxes[assign]
        h, w = pr[:, 2] - pr[:, 0], pr[:, 3] - pr[:, 1]
        cy, cx = pr[:, 0] + f32(0.5) * h, pr[:, 1] + f32(0.5) * w
        gh, gw = g[:, 2] - g[:, 0], g[:, 3] - g[:, 1]
        gcy, gcx = g[:, 0] + f32(0.5) * gh, g[:, 1] + f32(0.5) * gw
        d = np.stack([(gcy - cy) / h, (gcx - cx) / w, GL.log_f32(gh / h), GL.log_f32(gw / w)], axis=1).astype(f32)
        deltas[:pc] = d / np.asarray(bbox_std_dev, f32)
        rois[:pc] = pr
        cls[:pc] = gt_class_ids[assign]
        for t in range(pc):
            box = pr[t]
            if use_mini_mask:
                gy1, gx1, gy2, gx2 = g[t]
                box = np.array([(box[0] - gy1) / (gy2 - gy1), (box[1] - gx1) / (gx2 - gx1),
                                (box[2] - gy1) / (gy2 - gy1), (box[3] - gx1) / (gx2 - gx1)], f32)
            m = _crop_and_resize_1ch(gt_masks[:, :, assign[t]].astype(f32), box, mh, mw)
            masks[t] = np.round(m)            # tf.round: half to even, as numpy
    rois[pc:pc + len(neg_ix)] = proposals[neg_ix]
    return rois, cls, deltas, masks, (pc, len(neg_ix))


# --------------------------------------------------------------------------------------------------------------
# losses (torch float32, differentiable)
# --------------------------------------------------------------------------------------------------------------

def smooth_l1(y_true, y_pred):
    d = (y_true - y_pred).abs()
    lt = (d < 1.0).float()
    return lt * 0.5 * d ** 2 + (1 - lt) * (d - 0.5)


def losses(rpn_match, rpn_bbox_t, rpn_class_logits, rpn_bbox, target_class_ids, target_bbox, target_mask,
           mrcnn_class_logits, mrcnn_bbox, mrcnn_mask, active_class_ids):
    """The five loss graphs (mrcnn/model.py:1111-1270) -> dict of scalars.  Tensors are torch float32 / int64."""
    out = {}
    match = rpn_match.reshape(rpn_match.shape[0], -1)
    idx = torch.nonzero(match != 0)
    if len(idx):
        lg = rpn_class_logits[idx[:, 0], idx[:, 1]]
        out["rpn_class_loss"] = F.cross_entropy(lg, (match[idx[:, 0], idx[:, 1]] == 1).long())
    else:
        out["rpn_class_loss"] = torch.zeros(())
    pidx = torch.nonzero(match == 1)
    if len(pidx):
        pred = rpn_bbox[pidx[:, 0], pidx[:, 1]]
        counts = (match == 1).sum(1)
        tgt = torch.cat([rpn_bbox_t[b, :int(counts[b])] for b in range(match.shape[0])], 0).float()
        out["rpn_bbox_loss"] = smooth_l1(tgt, pred).mean()
    else:
        out["rpn_bbox_loss"] = torch.zeros(())
    tci = target_class_ids.long()
    nc = mrcnn_class_logits.shape[-1]
    ce = F.cross_entropy(mrcnn_class_logits.reshape(-1, nc), tci.reshape(-1), reduction="none")
    pred_active = active_class_ids[0].float()[mrcnn_class_logits.argmax(-1).reshape(-1)]
    out["mrcnn_class_loss"] = (ce * pred_active).sum() / pred_active.sum()
    ix = torch.nonzero(tci.reshape(-1) > 0)[:, 0]
    if len(ix):
        cls = tci.reshape(-1)[ix]
        out["mrcnn_bbox_loss"] = smooth_l1(target_bbox.reshape(-1, 4)[ix], mrcnn_bbox.reshape(-1, nc, 4)[ix, cls]).mean()
        y_true = target_mask.reshape((-1,) + tuple(target_mask.shape[2:]))[ix]
        y_pred = mrcnn_mask.reshape((-1,) + tuple(mrcnn_mask.shape[2:]))[ix, :, :, cls]
        o = y_pred.clamp(1e-7, 1 - 1e-7)                 # K.binary_crossentropy (Keras 2.2.4, TF backend)
        z = torch.log(o / (1 - o))
        out["mrcnn_mask_loss"] = (torch.clamp(z, min=0) - z * y_true + torch.log1p(torch.exp(-z.abs()))).mean()
    else:
        out["mrcnn_bbox_loss"] = torch.zeros(())
        out["mrcnn_mask_loss"] = torch.zeros(())
    return out


# --------------------------------------------------------------------------------------------------------------
# differentiable fp32 network
# --------------------------------------------------------------------------------------------------------------

def crop_and_resize(feat, boxes, pool):
    """feat [H,W,C] torch float32, boxes [n,4] (no gradient) -> [n,pool,pool,C]; differentiable in feat.
    Same coordinates as oracle/graph_layers.py crop_and_resize (float32)."""
    H, W, _ = feat.shape
    outs = []
    for bx in boxes:
        y1, x1, y2, x2 = [f32(v) for v in bx]
        Hm1, Wm1 = f32(H - 1), f32(W - 1)
        sy = (y2 - y1) * Hm1 / f32(pool - 1) if pool > 1 else f32(0)
        sx = (x2 - x1) * Wm1 / f32(pool - 1) if pool > 1 else f32(0)
        ys = np.array([y1 * Hm1 + f32(i) * sy for i in range(pool)], f32)
        xs = np.array([x1 * Wm1 + f32(i) * sx for i in range(pool)], f32)
        vy, vx = (ys >= 0) & (ys <= Hm1), (xs >= 0) & (xs <= Wm1)
        t, l = np.floor(np.where(vy, ys, 0)).astype(np.int64), np.floor(np.where(vx, xs, 0)).astype(np.int64)
        b, r = np.ceil(np.where(vy, ys, 0)).astype(np.int64), np.ceil(np.where(vx, xs, 0)).astype(np.int64)
        ly = torch.from_numpy((np.where(vy, ys, 0) - t).astype(f32)).view(-1, 1, 1)
        lx = torch.from_numpy((np.where(vx, xs, 0) - l).astype(f32)).view(1, -1, 1)
        tl, tr = feat[t][:, l], feat[t][:, r]
        bl, br = feat[b][:, l], feat[b][:, r]
        top = tl + (tr - tl) * lx
        bot = bl + (br - bl) * lx
        val = top + (bot - top) * ly
        valid = torch.from_numpy((vy[:, None] & vx[None, :]).astype(f32)).unsqueeze(-1)
        outs.append(val * valid)
    return torch.stack(outs) if outs else feat.new_zeros((0, pool, pool, feat.shape[2]))


class TrainNet(object):
    """fp32 torch-CPU training graph; weights {layer: [Keras-layout arrays]} become leaf tensors with gradients in
    self.p[(layer, role)] (role: kernel bias gamma beta; BN moving statistics are constants, TRAIN_BN=False)."""

    def __init__(self, weights, config):
        self.cfg = config
        self.p, self.stats = {}, {}
        for name, arrs in weights.items():
            if len(arrs) == 4:
                self.p[(name, "gamma")] = torch.tensor(arrs[0], dtype=torch.float32, requires_grad=True)
                self.p[(name, "beta")] = torch.tensor(arrs[1], dtype=torch.float32, requires_grad=True)
                self.stats[name] = (torch.tensor(arrs[2]), torch.tensor(arrs[3]))
            else:
                self.p[(name, "kernel")] = torch.tensor(arrs[0], dtype=torch.float32, requires_grad=True)
                self.p[(name, "bias")] = torch.tensor(arrs[1], dtype=torch.float32, requires_grad=True)

    def conv(self, x, name, bn=None, relu=False, stride=1, pad=0, residual=None):
        k = self.p[(name, "kernel")].permute(3, 2, 0, 1)
        y = F.conv2d(x, k, self.p[(name, "bias")], stride=stride, padding=pad)
        if bn is not None:
            mu, var = self.stats[bn]
            y = self.p[(bn, "gamma")].view(1, -1, 1, 1) * (y - mu.view(1, -1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1) + BN_EPS) \
                + self.p[(bn, "beta")].view(1, -1, 1, 1)
        if residual is not None:
            y = y + residual
        return F.relu(y) if relu else y

    def _block(self, x, stage, blk, first, stride):
        base, bnb = "res%d%s_branch" % (stage, blk), "bn%d%s_branch" % (stage, blk)
        y = self.conv(x, base + "2a", bnb + "2a", True, stride)
        y = self.conv(y, base + "2b", bnb + "2b", True, pad=1)
        sc = self.conv(x, base + "1", bnb + "1", stride=stride) if first else x
        return self.conv(y, base + "2c", bnb + "2c", True, residual=sc)

    def backbone_fpn(self, images):
        x = torch.as_tensor(images, dtype=torch.float32).permute(0, 3, 1, 2)
        x = self.conv(F.pad(x, (3, 3, 3, 3)), "conv1", "bn_conv1", True, 2)
        x = F.max_pool2d(F.pad(x, (0, 1, 0, 1), value=float("-inf")), 3, 2)
        C = {}
        for stage, n in ((2, 3), (3, 4), (4, 23), (5, 3)):
            for i in range(n):
                x = self._block(x, stage, chr(97 + i), i == 0, 2 if (i == 0 and stage > 2) else 1)
            C[stage] = x
        up = lambda t: F.interpolate(t, scale_factor=2, mode="nearest")
        p5 = self.conv(C[5], "fpn_c5p5")
        p4 = self.conv(C[4], "fpn_c4p4", residual=up(p5))
        p3 = self.conv(C[3], "fpn_c3p3", residual=up(p4))
        p2 = self.conv(C[2], "fpn_c2p2", residual=up(p3))
        P = {2: self.conv(p2, "fpn_p2", pad=1), 3: self.conv(p3, "fpn_p3", pad=1), 4: self.conv(p4, "fpn_p4", pad=1),
             5: self.conv(p5, "fpn_p5", pad=1)}
        P[6] = F.max_pool2d(P[5], 1, 2)
        return P

    def rpn(self, P):
        lg, bx = [], []
        for l in (2, 3, 4, 5, 6):
            s = self.conv(P[l], "rpn_conv_shared", relu=True, pad=1)
            c = self.conv(s, "rpn_class_raw").permute(0, 2, 3, 1)
            d = self.conv(s, "rpn_bbox_pred").permute(0, 2, 3, 1)
            lg.append(c.reshape(c.shape[0], -1, 2))
            bx.append(d.reshape(d.shape[0], -1, 4))
        lg = torch.cat(lg, 1)
        return lg, torch.softmax(lg, -1), torch.cat(bx, 1)

    def roi_align(self, rois, P, pool):
        """rois [B,T,4] numpy -> [B*T,C,pool,pool] (PyramidROIAlign, levels from oracle/graph_layers.py)"""
        S = float(self.cfg.IMAGE_SHAPE[0])
        out = []
        for b in range(rois.shape[0]):
            levels = GL.roi_levels(rois[b], S * S)
            feats = {l: P[l][b].permute(1, 2, 0) for l in (2, 3, 4, 5)}
            for t in range(rois.shape[1]):
                out.append(crop_and_resize(feats[int(levels[t])], rois[b, t:t + 1], pool)[0])
        return torch.stack(out).permute(0, 3, 1, 2)

    def class_head(self, rois, P):
        cfg = self.cfg
        x = self.roi_align(rois, P, int(cfg.POOL_SIZE))
        x = self.conv(x, "mrcnn_class_conv1", "mrcnn_class_bn1", True)
        x = self.conv(x, "mrcnn_class_conv2", "mrcnn_class_bn2", True)
        sh = x.reshape(x.shape[0], -1)
        logits = sh @ self.p[("mrcnn_class_logits", "kernel")] + self.p[("mrcnn_class_logits", "bias")]
        bbox = sh @ self.p[("mrcnn_bbox_fc", "kernel")] + self.p[("mrcnn_bbox_fc", "bias")]
        B, T = rois.shape[:2]
        nc = int(cfg.NUM_CLASSES)
        return logits.view(B, T, nc), bbox.view(B, T, nc, 4)

    def mask_head(self, rois, P):
        cfg = self.cfg
        x = self.roi_align(rois, P, int(cfg.MASK_POOL_SIZE))
        for i in range(1, 5):
            x = self.conv(x, "mrcnn_mask_conv%d" % i, "mrcnn_mask_bn%d" % i, True, pad=1)
        k = self.p[("mrcnn_mask_deconv", "kernel")].permute(3, 2, 0, 1)
        x = F.relu(F.conv_transpose2d(x, k, self.p[("mrcnn_mask_deconv", "bias")], stride=2))
        x = torch.sigmoid(self.conv(x, "mrcnn_mask"))
        B, T = rois.shape[:2]
        return x.permute(0, 2, 3, 1).reshape(B, T, x.shape[2], x.shape[3], -1)


def sgd_step(weights, grads, velocity, lr, momentum, clipnorm, weight_decay, world=1):
    """dicts keyed (layer, role) of float64/float32 numpy arrays, updated in place -> global norm before clipping.
    g = grad/world + 2*wd*w/size(w) for kernels and biases; g *= clipnorm / max(norm, clipnorm); v = m*v - lr*g; w += v."""
    g = {}
    for k, w in weights.items():
        reg = 2.0 * weight_decay / w.size if k[1] in ("kernel", "bias") else 0.0
        g[k] = grads[k].astype(np.float64) / world + reg * w.astype(np.float64)
    norm = float(np.sqrt(sum(float((v ** 2).sum()) for v in g.values())))
    scale = clipnorm / max(norm, clipnorm) if clipnorm and clipnorm > 0 else 1.0
    for k in weights:
        velocity[k] = momentum * velocity[k] - lr * scale * g[k]
        weights[k] = weights[k] + velocity[k]
    return norm
