#!/usr/bin/env python
"""e2e pipeline variants: where do the ~4 ms between device-only and pipelined host-in/host-out steps go?"""
import os, sys, time, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "caesar-mrcnn_b200"))
import numpy as np, torch, synth
from mrcnn import model as modellib, _native
from mrcnn.config import Config
B, S = 64, 256
class C(Config):
    NAME = "probe"; GPU_COUNT = 1; IMAGES_PER_GPU = B; NUM_CLASSES = 4; IMAGE_MIN_DIM = S; IMAGE_MAX_DIM = S
    RPN_ANCHOR_SCALES = (4, 8, 16, 32, 64); MEAN_PIXEL = np.array([0, 0, 0]); DETECTION_MIN_CONFIDENCE = 0
m = modellib.MaskRCNN("inference", C(), "/tmp/x"); m.set_weights(synth.make_random_weights(0, 4))
maps = torch.from_numpy(synth.radio_maps(B, S)).pin_memory()
dmaps = maps.cuda()
m.reserve_result_buffers(4, S, S)

def pipe(n, src, skip_masks=False):
    prev = None
    for i in range(n):
        if skip_masks:
            orig = m._result_buffers
            def rb(H0, W0, orig=orig):
                b = list(orig(H0, W0)); return tuple(b)
            h = m.detect_maps_async(src)
        else:
            h = m.detect_maps_async(src)
        if prev is not None: prev.result()
        prev = h
    prev.result()

def T(f, n=10):
    f(3); torch.cuda.synchronize(); t = time.perf_counter(); f(n); torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e3

print("device only                      %.2f ms/step" % T(lambda n: [m.detect_maps(dmaps, device_only=True) for _ in range(n)]))
print("pipeline host in / host out      %.2f ms/step" % T(lambda n: pipe(n, maps)))
print("pipeline device in / host out    %.2f ms/step" % T(lambda n: pipe(n, dmaps)))
# device-only compute with an unrelated 420 MB D2H copy in flight on another stream
big = torch.empty((B * S * S * 100,), dtype=torch.uint8, device="cuda"); hbig = torch.empty_like(big, device="cpu").pin_memory()
side = torch.cuda.Stream()
def with_copy(n):
    for _ in range(n):
        with torch.cuda.stream(side):
            hbig.copy_(big, non_blocking=True)
        m.detect_maps(dmaps, device_only=True)
    side.synchronize()
print("device only + concurrent 420MB D2H %.2f ms/step" % T(with_copy))
def with_h2d(n):
    for _ in range(n):
        with torch.cuda.stream(side):
            big.copy_(hbig, non_blocking=True)
        m.detect_maps(dmaps, device_only=True)
    side.synchronize()
print("device only + concurrent 420MB H2D %.2f ms/step" % T(with_h2d))
