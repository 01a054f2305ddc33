#!/usr/bin/env python
"""Latency probe: back-to-back device-only steps and synchronous detect_maps calls at a given batch size."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "caesar-mrcnn_b200"))
import numpy as np, torch, synth
from mrcnn import model as modellib
from mrcnn.config import Config
B, S = int(sys.argv[1]) if len(sys.argv) > 1 else 1, 256
class C(Config):
    NAME = "probe"; GPU_COUNT = 1; IMAGES_PER_GPU = B; NUM_CLASSES = 4; IMAGE_MIN_DIM = S; IMAGE_MAX_DIM = S
    RPN_ANCHOR_SCALES = (4, 8, 16, 32, 64); MEAN_PIXEL = np.array([0, 0, 0]); DETECTION_MIN_CONFIDENCE = 0
m = modellib.MaskRCNN("inference", C(), "/tmp/x"); m.set_weights(synth.make_random_weights(0, 4))
maps = torch.from_numpy(synth.radio_maps(B, S)).cuda()
def run(n):
    for i in range(n): m.detect_maps(maps, device_only=True, _async=True)
    m.wait()
best = 1e9
for rep in range(4):
    run(10); torch.cuda.synchronize(); t = time.perf_counter(); run(100); torch.cuda.synchronize(); best = min(best, (time.perf_counter() - t) / 100 * 1e3)
hm = maps.cpu().pin_memory()
lat = []
for i in range(30):
    t = time.perf_counter(); r = m.detect_maps(hm); lat.append((time.perf_counter() - t) * 1e3)
print("B=%d GRAPH=%s PDL=%s CHAIN=%s: %.3f ms/batch back-to-back, sync call latency median %.3f ms" % (
    B, os.environ.get("MRCNN_B200_GRAPH", "default"), os.environ.get("MRCNN_B200_PDL", "1"), os.environ.get("MRCNN_B200_CHAIN", "0"),
    best, sorted(lat)[len(lat) // 2]))
