#!/usr/bin/env python
"""Short profiling workload: two detect_maps steps at B=64, S=256 (the first is warm-up).
Used under ncu with -k filters; a number printed by a run under ncu is never a bench value."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "caesar-mrcnn_b200"))
import numpy as np
import torch
import synth
from mrcnn import model as modellib
from mrcnn.config import Config

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 2
S = 256


class C(Config):
    NAME = "prof"
    GPU_COUNT = 1
    IMAGES_PER_GPU = B
    NUM_CLASSES = 4
    IMAGE_MIN_DIM = S
    IMAGE_MAX_DIM = S
    RPN_ANCHOR_SCALES = (4, 8, 16, 32, 64)
    MEAN_PIXEL = np.array([0, 0, 0])
    DETECTION_MIN_CONFIDENCE = 0


m = modellib.MaskRCNN("inference", C(), "/tmp/x")
m.set_weights(synth.make_random_weights(0, 4))
maps = torch.from_numpy(synth.radio_maps(B, S)).cuda()
from mrcnn import _native
lib = _native.lib()
m.detect_maps(maps, device_only=True)
n0 = lib.mrcnn_kernel_launch_count()
for _ in range(STEPS - 1):
    m.detect_maps(maps, device_only=True)
torch.cuda.synchronize()
print("profile_run ok launches_per_step=%d" % ((lib.mrcnn_kernel_launch_count() - n0) // max(1, STEPS - 1)))
