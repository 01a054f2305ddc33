"""ORACLE (test infrastructure) — torch-CPU float32 restatement of the dense part of the graph.

  resnet101 backbone   <- mrcnn/model.py:99-210 (identity_block, conv_block, resnet_graph)
  FPN wiring           <- mrcnn/model.py:2003-2026
  RPN                  <- mrcnn/model.py:916-979, per-level application + concat :2040-2055
  class / bbox head    <- mrcnn/model.py:986-1039 (fpn_classifier_graph)
  mask head            <- mrcnn/model.py:1042-1091 (build_fpn_mask_graph)
  whole inference graph<- mrcnn/model.py:2133-2159 (outputs in the same order)

Keras/TF semantics restated: BN inference with eps=1e-3, TF SAME padding
(max-pool 3x3 s2 pads bottom/right only), nearest 2x upsampling, Conv2DTranspose 2x2 s2 VALID with
kernel [kh,kw,Cout,Cin], softmax over the last axis.  Weight layouts follow Appendix B.

Two arithmetic modes:
  emulate_bf16=False : plain fp32 (the "reference" numerics).
  emulate_bf16=True  : what the CUDA engine computes, restated on the CPU — weights and every
                       stored activation rounded to bf16, fp32 accumulation, BN folded as
                       y = acc*s + t.  Used for tight kernel-vs-oracle comparisons.
parity: UNPINNED against TF/Keras (cannot run here); torch.nn.functional is the independent check.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import graph_layers as GL

BN_EPS = 1e-3


# layer inventory + seeded random weights live in synth.py (neutral workload generator shared with
# bench.py); re-exported here for the tests
from synth import layer_specs, make_random_weights  # noqa: E402,F401


# --------------------------------------------------------------------------------------------
# the network
# --------------------------------------------------------------------------------------------

def _bf16(t):
    return t.to(torch.bfloat16).to(torch.float32)


SECTIONS = ("backbone", "fpn", "rpn", "class", "mask")


class OracleNet:
    def __init__(self, weights, num_classes=4, emulate_bf16=False, threads=None, trunk_fp32=False):
        """emulate_bf16: False (fp32 everywhere), True (the engine's arithmetic everywhere) or a set of SECTIONS that
        use the engine's arithmetic while the others stay fp32 (precision studies, tools/e2e_parity_cpu.py).
        trunk_fp32: study variant — the ResNet residual trunk (block outputs) is stored in fp32 and only rounded to
        bf16 where it is consumed as a GEMM operand."""
        self.w = weights
        self.nc = num_classes
        self._emu_sections = set(SECTIONS) if emulate_bf16 is True else set(emulate_bf16 or ())
        self.emu = bool(self._emu_sections)
        self.trunk_fp32 = trunk_fp32
        if threads:
            torch.set_num_threads(threads)
        self._cache = {}

    # -- parameter helpers ------------------------------------------------------------------
    def _section(self, name):
        self.emu = name in self._emu_sections

    def _kernel(self, name):
        key = (name, self.emu)
        if key not in self._cache:
            k = torch.from_numpy(np.ascontiguousarray(self.w[name][0]))
            if k.ndim == 4:
                k = k.permute(3, 2, 0, 1).contiguous()    # [kh,kw,Cin,Cout] -> [Cout,Cin,kh,kw]
            if self.emu:
                k = _bf16(k)
            self._cache[key] = k
        return self._cache[key]

    def _bias(self, name):
        return torch.from_numpy(self.w[name][1])

    def _affine(self, conv, bn):
        """(s, t) with y = acc*s + t, acc = bias-free convolution."""
        b = self._bias(conv).double()
        if bn is None:
            return torch.ones_like(b).float(), b.float()
        g, be, mu, var = [torch.from_numpy(a).double() for a in self.w[bn]]
        s = g / torch.sqrt(var + BN_EPS)
        t = (b - mu) * s + be
        return s.float(), t.float()

    def _finish(self, acc, conv, bn, relu, residual=None, keep_f32=False):
        if self.emu:
            s, t = self._affine(conv, bn)
            y = acc * s.view(1, -1, 1, 1) + t.view(1, -1, 1, 1)
        else:
            y = acc + self._bias(conv).view(1, -1, 1, 1)
            if bn is not None:
                g, be, mu, var = [torch.from_numpy(a).view(1, -1, 1, 1) for a in self.w[bn]]
                y = g * (y - mu) / torch.sqrt(var + BN_EPS) + be
        if residual is not None:
            y = y + residual
        if relu:
            y = F.relu(y)
        return _bf16(y) if (self.emu and not keep_f32) else y

    def conv(self, x, name, bn=None, relu=False, stride=1, pad=0, residual=None, keep_f32=False):
        if self.emu:
            x = _bf16(x)          # GEMM operands are bf16 (a no-op for activations that were stored as bf16)
        acc = F.conv2d(x, self._kernel(name), None, stride=stride, padding=pad)
        return self._finish(acc, name, bn, relu, residual, keep_f32)

    # -- backbone + FPN ---------------------------------------------------------------------
    def _block(self, x, stage, blk, first, stride):
        base = "res%d%s_branch" % (stage, blk)
        bnb = "bn%d%s_branch" % (stage, blk)
        y = self.conv(x, base + "2a", bnb + "2a", relu=True, stride=stride)
        y = self.conv(y, base + "2b", bnb + "2b", relu=True, pad=1)
        if first:
            sc = self.conv(x, base + "1", bnb + "1", relu=False, stride=stride, keep_f32=self.trunk_fp32)
        else:
            sc = x
        return self.conv(y, base + "2c", bnb + "2c", relu=True, residual=sc, keep_f32=self.trunk_fp32)

    def backbone_fpn(self, molded):
        """molded [B,S,S,3] float32 NHWC -> dict with C2..C5, P2..P6 (NCHW torch tensors)."""
        self._section("backbone")
        x = torch.from_numpy(np.ascontiguousarray(molded, dtype=np.float32)).permute(0, 3, 1, 2)
        if self.emu:
            x = _bf16(x)
        x = self.conv(F.pad(x, (3, 3, 3, 3)), "conv1", "bn_conv1", relu=True, stride=2)
        # MaxPooling2D(3, strides=2, padding='same'): TF pads (total = max((out-1)*2+3-in, 0))
        # before = total//2, remainder after -> for even sizes 0 before / 1 after.
        H, W = x.shape[2:]
        ph = max((-(-H // 2) - 1) * 2 + 3 - H, 0)
        pw = max((-(-W // 2) - 1) * 2 + 3 - W, 0)
        x = F.pad(x, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2), value=float("-inf"))
        x = F.max_pool2d(x, 3, 2)
        feats = {}
        for stage, blocks in ((2, "abc"), (3, "abcd"),
                              (4, "a" + "".join(chr(98 + i) for i in range(22))), (5, "abc")):
            for blk in blocks:
                first = blk == "a"
                stride = 2 if (first and stage > 2) else 1
                x = self._block(x, stage, blk, first, stride)
            feats["C%d" % stage] = x
        self._section("fpn")
        p5 = self.conv(feats["C5"], "fpn_c5p5")
        p4 = self.conv(feats["C4"], "fpn_c4p4", residual=F.interpolate(p5, scale_factor=2, mode="nearest"))
        p3 = self.conv(feats["C3"], "fpn_c3p3", residual=F.interpolate(p4, scale_factor=2, mode="nearest"))
        p2 = self.conv(feats["C2"], "fpn_c2p2", residual=F.interpolate(p3, scale_factor=2, mode="nearest"))
        feats["P2"] = self.conv(p2, "fpn_p2", pad=1)
        feats["P3"] = self.conv(p3, "fpn_p3", pad=1)
        feats["P4"] = self.conv(p4, "fpn_p4", pad=1)
        feats["P5"] = self.conv(p5, "fpn_p5", pad=1)
        feats["P6"] = feats["P5"][:, :, ::2, ::2].contiguous()
        return feats

    # -- RPN --------------------------------------------------------------------------------
    def rpn(self, feats):
        """-> rpn_class [B,A,2], rpn_bbox [B,A,4] float32 numpy (level-major anchor order)."""
        self._section("rpn")
        cls, box = [], []
        for lv in ("P2", "P3", "P4", "P5", "P6"):
            shared = self.conv(feats[lv], "rpn_conv_shared", relu=True, pad=1)
            a = F.conv2d(shared, self._kernel("rpn_class_raw"), None) \
                + self._bias("rpn_class_raw").view(1, -1, 1, 1)
            b = F.conv2d(shared, self._kernel("rpn_bbox_pred"), None) \
                + self._bias("rpn_bbox_pred").view(1, -1, 1, 1)
            B = a.shape[0]
            cls.append(a.permute(0, 2, 3, 1).reshape(B, -1, 2))
            box.append(b.permute(0, 2, 3, 1).reshape(B, -1, 4))
        logits = torch.cat(cls, dim=1)
        probs = torch.softmax(logits, dim=-1)
        return probs.numpy(), torch.cat(box, dim=1).numpy()

    # -- heads ------------------------------------------------------------------------------
    def class_head(self, pooled):
        """pooled [B,N,7,7,C] float32 -> mrcnn_class [B,N,NC], mrcnn_bbox [B,N,NC,4]."""
        self._section("class")
        B, N = pooled.shape[:2]
        x = torch.from_numpy(np.ascontiguousarray(pooled, dtype=np.float32))
        x = x.reshape(B * N, *pooled.shape[2:]).permute(0, 3, 1, 2)
        if self.emu:
            x = _bf16(x)
        x = self.conv(x, "mrcnn_class_conv1", "mrcnn_class_bn1", relu=True)
        x = self.conv(x, "mrcnn_class_conv2", "mrcnn_class_bn2", relu=True)
        shared = x.reshape(B * N, -1)
        wl = torch.from_numpy(self.w["mrcnn_class_logits"][0])
        wb = torch.from_numpy(self.w["mrcnn_bbox_fc"][0])
        if self.emu:
            wl, wb = _bf16(wl), _bf16(wb)
        logits = shared @ wl + torch.from_numpy(self.w["mrcnn_class_logits"][1])
        bbox = shared @ wb + torch.from_numpy(self.w["mrcnn_bbox_fc"][1])
        probs = torch.softmax(logits, dim=-1)
        return (probs.reshape(B, N, self.nc).numpy(),
                bbox.reshape(B, N, self.nc, 4).numpy())

    def mask_head(self, pooled):
        """pooled [B,N,14,14,C] float32 -> mrcnn_mask [B,N,28,28,NC]."""
        self._section("mask")
        B, N = pooled.shape[:2]
        x = torch.from_numpy(np.ascontiguousarray(pooled, dtype=np.float32))
        x = x.reshape(B * N, *pooled.shape[2:]).permute(0, 3, 1, 2)
        if self.emu:
            x = _bf16(x)
        for i in range(1, 5):
            x = self.conv(x, "mrcnn_mask_conv%d" % i, "mrcnn_mask_bn%d" % i, relu=True, pad=1)
        # Conv2DTranspose kernel [kh,kw,Cout,Cin] -> torch conv_transpose2d weight [Cin,Cout,kh,kw]
        wd = torch.from_numpy(np.ascontiguousarray(self.w["mrcnn_mask_deconv"][0])).permute(3, 2, 0, 1)
        if self.emu:
            wd = _bf16(wd)
        x = F.conv_transpose2d(x, wd.contiguous(), None, stride=2)
        x = F.relu(x + self._bias("mrcnn_mask_deconv").view(1, -1, 1, 1))
        if self.emu:
            x = _bf16(x)
        x = F.conv2d(x, self._kernel("mrcnn_mask"), None) + self._bias("mrcnn_mask").view(1, -1, 1, 1)
        x = torch.sigmoid(x)
        return x.permute(0, 2, 3, 1).reshape(B, N, x.shape[2], x.shape[3], self.nc).numpy()

    # -- whole graph ------------------------------------------------------------------------
    def predict(self, molded, image_metas, anchors, cfg):
        """keras_model.predict([molded_images, image_metas, anchors]) for mode='inference'.
        cfg: dict with PRE_NMS_LIMIT, POST_NMS_ROIS_INFERENCE, RPN_NMS_THRESHOLD, RPN_BBOX_STD_DEV,
        BBOX_STD_DEV, DETECTION_MIN_CONFIDENCE, DETECTION_NMS_THRESHOLD, DETECTION_MAX_INSTANCES,
        POOL_SIZE, MASK_POOL_SIZE.  Returns a dict keyed like model.py:2156-2158 (+ taps)."""
        feats = self.backbone_fpn(molded)
        rpn_class, rpn_bbox = self.rpn(feats)
        rpn_rois = GL.proposal_layer(rpn_class, rpn_bbox, anchors,
                                     pre_nms_limit=cfg["PRE_NMS_LIMIT"],
                                     proposal_count=cfg["POST_NMS_ROIS_INFERENCE"],
                                     nms_threshold=cfg["RPN_NMS_THRESHOLD"],
                                     rpn_bbox_std_dev=cfg["RPN_BBOX_STD_DEV"])
        fmaps = [feats[k].permute(0, 2, 3, 1).contiguous().numpy() for k in ("P2", "P3", "P4", "P5")]
        image_shape = np.asarray(image_metas)[0, 4:7]
        p = cfg["POOL_SIZE"]
        pooled = GL.pyramid_roi_align(rpn_rois, image_shape, fmaps, (p, p))
        mrcnn_class, mrcnn_bbox = self.class_head(pooled)
        detections = GL.detection_layer(rpn_rois, mrcnn_class, mrcnn_bbox, image_metas,
                                        bbox_std_dev=cfg["BBOX_STD_DEV"],
                                        min_confidence=cfg["DETECTION_MIN_CONFIDENCE"],
                                        nms_threshold=cfg["DETECTION_NMS_THRESHOLD"],
                                        max_instances=cfg["DETECTION_MAX_INSTANCES"])
        mp = cfg["MASK_POOL_SIZE"]
        pooled_m = GL.pyramid_roi_align(detections[..., :4], image_shape, fmaps, (mp, mp))
        mrcnn_mask = self.mask_head(pooled_m)
        return {"detections": detections, "mrcnn_class": mrcnn_class, "mrcnn_bbox": mrcnn_bbox,
                "mrcnn_mask": mrcnn_mask, "rpn_rois": rpn_rois, "rpn_class": rpn_class,
                "rpn_bbox": rpn_bbox, "fmaps": fmaps, "P6": feats["P6"].permute(0, 2, 3, 1).numpy(),
                "pooled": pooled, "pooled_mask": pooled_m}
