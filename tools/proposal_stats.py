#!/usr/bin/env python
"""How deep into the 6000 candidates does the ProposalLayer NMS go on the bench workload?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "caesar-mrcnn_b200"))
import numpy as np, torch, synth
from mrcnn import model as modellib
from mrcnn.config import Config
B, S = 64, 256
class C(Config):
    NAME = "probe"; GPU_COUNT = 1; IMAGES_PER_GPU = B; NUM_CLASSES = 4; IMAGE_MIN_DIM = S; IMAGE_MAX_DIM = S
    RPN_ANCHOR_SCALES = (4, 8, 16, 32, 64); MEAN_PIXEL = np.array([0, 0, 0]); DETECTION_MIN_CONFIDENCE = 0
m = modellib.MaskRCNN("inference", C(), "/tmp/x"); m.set_weights(synth.make_random_weights(0, 4))
maps = torch.from_numpy(synth.radio_maps(B, S)).cuda()
m.detect_maps(maps, device_only=True)
keep = m.read_tensor("keep_idx"); cnt = m.read_tensor("keep_count")
sc = m.read_tensor("rpn_class")[:, :, 1]
top = -np.sort(-sc, axis=1)[:, :6000]
ties = (top[:, 1:] == top[:, :-1]).sum(axis=1)
print("keep_count   min/mean/max:", cnt.min(), cnt.mean(), cnt.max())
print("deepest kept candidate (pop position) min/mean/max:", keep.max(axis=1).min(), keep.max(axis=1).mean(), keep.max(axis=1).max())
print("adjacent equal scores among the top 6000: min/mean/max:", ties.min(), ties.mean(), ties.max())
print("score range of top-6000: %.6f .. %.6f" % (top[:, -1].min(), top[:, 0].max()))
lv = m.read_tensor("roi_levels"); print("roi levels histogram:", np.bincount(lv.ravel(), minlength=6)[2:])
