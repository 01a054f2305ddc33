#!/usr/bin/env python
"""Per-launch timing table of one detect step (CUDA events around every launch, B=64, S=256).
Usage (on a GPU box): python tools/layer_table.py [batch] > gpurun_out/layer_table.txt"""
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
S = 256


class C(Config):
    NAME = "tbl"
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
m.set_profiling(True)
maps = torch.from_numpy(synth.radio_maps(B, S)).cuda()
acc = None
for it in range(5):
    m.detect_maps(maps, device_only=True)
    t = m.step_table()
    if it >= 2:
        acc = t if acc is None else [(a[0], a[1], a[2] + b[2], a[3]) for a, b in zip(acc, t)]
n = 3
tot = sum(a[2] for a in acc) / n
print("%-34s %-12s %9s %9s %7s" % ("launch", "family", "ms", "TFLOP/s", "%step"))
for label, kind, ms, fl in acc:
    ms /= n
    print("%-34s %-12s %9.4f %9.1f %6.2f%%" % (label, kind, ms, fl / ms / 1e9 if fl else 0.0, 100 * ms / tot))
print("total %.3f ms" % tot)
