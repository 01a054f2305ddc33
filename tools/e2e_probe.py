#!/usr/bin/env python
"""Where does the end-to-end time go? (host-side timing of the public API phases)"""
import os, sys, time
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
maps = torch.from_numpy(synth.radio_maps(B, S)).pin_memory()
dmaps = maps.cuda()
def T(f, n=5):
    f(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e3
print("device_only           %.2f ms" % T(lambda: m.detect_maps(dmaps, device_only=True)))
print("host in, device out   %.2f ms" % T(lambda: m.detect_maps(maps, device_only=True)))
print("sync host in/out      %.2f ms" % T(lambda: m.detect_maps(maps)))
print("alloc result buffers  %.2f ms" % T(lambda: m._result_buffers(S, S)))
bufs = m._result_buffers(S, S)
print("views from buffers    %.2f ms" % T(lambda: m._results_from_buffers(bufs, B)))
def pipe(n=6):
    prev = None
    for i in range(n):
        h = m.detect_maps_async(maps)
        if prev is not None: prev.result()
        prev = h
    prev.result()
pipe(3); torch.cuda.synchronize(); t = time.perf_counter(); pipe(10); print("async pipeline        %.2f ms/step" % ((time.perf_counter() - t) / 10 * 1e3))
# queue two without touching results
t = time.perf_counter(); h1 = m.detect_maps_async(maps); t1 = time.perf_counter(); h2 = m.detect_maps_async(maps); t2 = time.perf_counter()
h1.result(); t3 = time.perf_counter(); h2.result(); t4 = time.perf_counter()
print("queue1 %.2f queue2 %.2f result1 %.2f result2 %.2f ms" % ((t1 - t) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3))
# bench pattern: keep the previous step's result dicts alive
keep = [None]; prev = None; ts = []
for i in range(12):
    t0 = time.perf_counter()
    h = m.detect_maps_async(maps)
    if prev is not None: keep[0] = prev.result()
    prev = h
    ts.append((time.perf_counter() - t0) * 1e3)
keep[0] = prev.result()
print("retaining results, per-step host ms:", " ".join("%.1f" % t for t in ts))
with torch.cuda.stream(m._stream):
    prev = None; ts = []
    for i in range(12):
        t0 = time.perf_counter()
        h = m.detect_maps_async(maps)
        if prev is not None: keep[0] = prev.result()
        prev = h
        ts.append((time.perf_counter() - t0) * 1e3)
    keep[0] = prev.result()
print("same under torch.cuda.stream(engine stream):", " ".join("%.1f" % t for t in ts))
