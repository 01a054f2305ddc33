#!/usr/bin/env python
"""CPU study (no GPU): final detections of the bf16-emulating oracle (what the engine computes) against the fp32
oracle on the same synthetic maps and weights -> the distribution the GPU test (tests/test_gpu_e2e_parity.py) must
reproduce.  python tools/e2e_parity_cpu.py [n_images] [out.json]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import parity_metrics as PM  # noqa: E402
import synth  # noqa: E402
from bench import _oracle_cfg  # noqa: E402
from oracle import host_ops as H, network as N  # noqa: E402

S = 256
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
weights = synth.make_random_weights(0, 4)
anchors = H.get_anchors((S, S, 3), (4, 8, 16, 32, 64))
VARIANTS = {"bf16": dict(emulate_bf16=True),                                   # the engine's arithmetic
            "bf16_trunk_fp32": dict(emulate_bf16=True, trunk_fp32=True),         # candidate fix: fp32 residual trunk
            "bf16_backbone_only": dict(emulate_bf16={"backbone"}),              # where does the error come from?
            "bf16_heads_only": dict(emulate_bf16={"fpn", "rpn", "class", "mask"})}
variant = os.environ.get("PARITY_VARIANT", "bf16")
nets = {"fp32": N.OracleNet(weights, 4, emulate_bf16=False), "bf16": N.OracleNet(weights, 4, **VARIANTS[variant])}
per_det, per_res, stage = [], [], []
for i in range(n):
    m = synth.radio_map(i, S)
    img = H.fits_to_rgb(m)
    molded, metas, windows = H.mold_inputs([img], min_dim=S, max_dim=S, min_scale=0, mode="square",
                                           mean_pixel=np.array([0, 0, 0]), num_classes=4)
    outs, res = {}, {}
    for k, net in nets.items():
        o = net.predict(molded, metas, anchors, _oracle_cfg())
        b, c, s, mk = H.unmold_detections(o["detections"][0], o["mrcnn_mask"][0], img.shape, (S, S, 3), windows[0])
        outs[k], res[k] = o, {"rois": b, "class_ids": c, "scores": s, "masks": mk}
    per_det.append(PM.match_detections_tensor(outs["bf16"]["detections"][0], outs["fp32"]["detections"][0], S))
    per_res.append(PM.match_results(res["bf16"], res["fp32"]))
    stage.append(PM.stage_errors(outs["bf16"], outs["fp32"]))
    print(i, PM.summarize(per_det[-1:]), flush=True)
out = {"variant": variant, "detections_tensor": PM.summarize(per_det), "unmolded": PM.summarize(per_res),
       "stage_max": {k: max(s[k] for s in stage) for k in stage[0]}}
print(json.dumps(out, indent=1))
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], "w"), indent=1)
