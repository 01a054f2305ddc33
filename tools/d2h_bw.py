#!/usr/bin/env python
"""Pinned host<->device copy bandwidth of the box (ceiling for the e2e number)."""
import torch
n = 420 * 1024 * 1024
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
for name, src, dst in (("d2h", d, h), ("h2d", h, d)):
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        dst.copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("%s %.1f MB in %.2f ms = %.1f GB/s" % (name, n / 1e6, ms, n / ms / 1e6))
