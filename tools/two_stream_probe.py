#!/usr/bin/env python
"""Do two engines on two streams (two batches in flight) beat one engine? Device-only steps, B=64, S=256."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "caesar-mrcnn_b200"))
import numpy as np, torch, synth
from mrcnn import model as modellib
from mrcnn.config import Config
B, S = int(sys.argv[1]) if len(sys.argv) > 1 else 64, 256
class C(Config):
    NAME = "probe"; GPU_COUNT = 1; IMAGES_PER_GPU = B; NUM_CLASSES = 4; IMAGE_MIN_DIM = S; IMAGE_MAX_DIM = S
    RPN_ANCHOR_SCALES = (4, 8, 16, 32, 64); MEAN_PIXEL = np.array([0, 0, 0]); DETECTION_MIN_CONFIDENCE = 0
w = synth.make_random_weights(0, 4)
ms = []
for k in range(2):
    m = modellib.MaskRCNN("inference", C(), "/tmp/x"); m.set_weights(w); ms.append(m)
maps = torch.from_numpy(synth.radio_maps(B, S)).cuda()
def run(models, n):
    for i in range(n):
        models[i % len(models)].detect_maps(maps, device_only=True, _async=True)
    for m in models: m.wait()
def T(models, n=20):
    run(models, 4); torch.cuda.synchronize(); t = time.perf_counter(); run(models, n); torch.cuda.synchronize()
    return (time.perf_counter() - t) / n * 1e3
print("one engine , async queue : %.3f ms/batch" % T(ms[:1]))
print("two engines, two streams : %.3f ms/batch" % T(ms))
print("one engine , async queue : %.3f ms/batch" % T(ms[:1]))
