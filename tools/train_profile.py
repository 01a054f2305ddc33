#!/usr/bin/env python
"""Kernel-level time split of one eager training step (torch.profiler / CUPTI): python tools/train_profile.py [out.json]"""
import collections
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "caesar-mrcnn_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import synth  # noqa: E402
import train_bench  # noqa: E402
from mrcnn import training  # noqa: E402

cfg = train_bench._config()
graph = training.TrainGraph(cfg, device=0, layers="all", seed=0)
graph.params.set_weights(synth.make_random_weights(0, 4))
trainer = training.Trainer(graph)
np.random.seed(1)
gen = training.data_generator(train_bench._dataset(0, 2 * cfg.BATCH_SIZE), cfg, shuffle=False, batch_size=cfg.BATCH_SIZE)
dev = graph.to_device(next(gen)[0])
for _ in range(3):
    trainer.train_step(dev)
torch.cuda.synchronize()
if os.environ.get("NO_TORCH_PROF") == "1":          # under ncu: just one more eager step (CUPTI cannot be shared)
    trainer.train_step(dev)
    torch.cuda.synchronize()
    print("one eager step done (no torch profiler)")
    sys.exit(0)
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    trainer.train_step(dev)
    torch.cuda.synchronize()
tot, cnt = collections.Counter(), collections.Counter()
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        name = re.sub(r"^void ", "", e.name.replace("(anonymous namespace)::", "").replace("<unnamed>::", ""))
        name = re.match(r"[\w:]+", name).group(0)[:60] if re.match(r"[\w:]+", name) else name[:60]
        tot[name] += e.device_time if hasattr(e, "device_time") else e.cuda_time
        cnt[name] += 1
total = sum(tot.values())
out = {"total_us": total, "kernels": sum(cnt.values()), "top": [{"name": k, "us": v, "launches": cnt[k]} for k, v in tot.most_common(25)]}
print("one eager step: %d kernels, %.2f ms of kernel time" % (out["kernels"], total / 1e3))
for t in out["top"]:
    print("%9.1f us %5d  %s" % (t["us"], t["launches"], t["name"]))
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
