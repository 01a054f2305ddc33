#!/usr/bin/env python
"""Host side of the result path: time mrcnn_host_expand_mask_bits (pixel-major mask bits -> [H,W,N] bool arrays) for one
64-image batch of 256x256 frames with 100 detections each, over thread counts; no GPU needed.
python tools/host_expand_bench.py [out.json]"""
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "caesar-mrcnn_b200"))
import numpy as np  # noqa: E402
from mrcnn import _native  # noqa: E402

lib = _native.lib()
B, npx, D = 64, 256 * 256, 100
dw = lib.mrcnn_mask_bits_words(D)
rng = np.random.default_rng(0)
bits = rng.integers(0, 2 ** 32, size=(B, npx, dw), dtype=np.uint64).astype(np.uint32)
counts = np.full((B,), D, dtype=np.int32)
sets = [np.zeros((B, npx * D), np.uint8) for _ in range(3)]          # rotate destinations like the result pool does
out = {"cpus": len(os.sched_getaffinity(0)), "default_threads": int(lib.mrcnn_host_threads()), "bytes_out": B * npx * D, "runs": []}
for threads in (1, 2, 4, 8, 16, 32):
    if threads > out["cpus"]:
        break
    ts = []
    for rep in range(7):
        dense = sets[rep % 3]
        dst = (ctypes.c_void_p * B)(*[dense[i].ctypes.data for i in range(B)])
        t0 = time.perf_counter()
        _native.check(lib.mrcnn_host_expand_mask_bits(bits.ctypes.data, B, npx, dw, counts.ctypes.data, dst, threads))
        ts.append(time.perf_counter() - t0)
    ms = 1e3 * min(ts[2:])
    out["runs"].append({"threads": threads, "ms": ms, "gb_per_s_written": B * npx * D / ms / 1e6})
    print("threads %2d: %.2f ms  (%.1f GB/s written)" % (threads, ms, B * npx * D / ms / 1e6), flush=True)
print(json.dumps(out))
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
