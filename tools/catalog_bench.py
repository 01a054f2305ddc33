#!/usr/bin/env python
"""End-to-end detect + Analyzer throughput: pinned host maps in, source catalogues (the reference's JSON objects
minus 'vertexes'; "pixels" as one int32 array per object) out, through `Analyzer.predict_maps_stream` (detect of batch k+1 overlaps the post-processing of
batch k; the [B,H,W,100] masks never leave the GPU).

Same workload as bench.py (BASELINE.json configs[1]: 64 synthetic maps at IMAGE_MAX_DIM=256, random weights). The
score threshold is set to the median detection score of the first batch, so that about half of the 100 raw
detections per image go through merging / selection (random weights give no score above the default 0.7).
Prints one JSON line; wall clock around K steps with the device drained on both sides."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "caesar-mrcnn_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--image-size", type=int, default=256)
    args = ap.parse_args()

    import torch
    import synth
    from mrcnn import model as modellib
    from mrcnn.analyze import Analyzer
    from mrcnn.config import Config

    B, S = args.batch, args.image_size

    class BenchConfig(Config):
        NAME = "rg-dataset"
        GPU_COUNT = 1
        IMAGES_PER_GPU = B
        NUM_CLASSES = 4
        CLASS_NAMES = ["bkg", "sidelobe", "source", "galaxy"]
        IMAGE_MIN_DIM = S
        IMAGE_MAX_DIM = S
        RPN_ANCHOR_SCALES = (4, 8, 16, 32, 64)
        MEAN_PIXEL = np.array([0, 0, 0])
        DETECTION_MIN_CONFIDENCE = 0
        RPN_NMS_THRESHOLD = 0.7

    cfg = BenchConfig()
    model = modellib.MaskRCNN(mode="inference", config=cfg, model_dir="/tmp/mrcnn_bench", device=0)
    model.set_weights(synth.make_random_weights(0, 4))
    base = synth.radio_maps(B, S, start=0)
    host_sets = []
    for k in range(4):
        arr = np.roll(base, k * 7, axis=0).copy()
        if k % 2:
            arr = arr[:, ::-1, :].copy()
        host_sets.append(torch.from_numpy(arr).pin_memory())

    first = model.detect_maps(host_sets[0])
    scores = np.concatenate([r["scores"] for r in first])
    an = Analyzer(model, cfg)
    an.score_thr = float(np.median(scores)) if len(scores) else 0.7
    del first

    def run(steps):
        n_obj = n_pix = 0
        for cats in an.predict_maps_stream(host_sets[i % 4] for i in range(steps)):
            for cat in cats:
                n_obj += len(cat["objs"])
                n_pix += sum(len(o["pixels"]) for o in cat["objs"])
        return n_obj, n_pix

    run(args.warmup)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n_obj, n_pix = run(args.steps)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0

    # serial variant for comparison (no overlap): detect, wait, analyse
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(args.steps):
        an.predict_maps(host_sets[i % 4])
    torch.cuda.synchronize()
    wall_serial = time.perf_counter() - t0

    # stage breakdown, serial calls (separate pass: every mark drains the device)
    an._timings = {}
    for i in range(6):
        an.predict_maps(host_sets[i % 4])
    stages = {k: round(v / 6 * 1e3, 2) for k, v in an._timings.items()}
    an._timings = None

    D = cfg.DETECTION_MAX_INSTANCES
    print(json.dumps({
        "metric": "catalogued_images_per_sec", "value": B * args.steps / wall, "unit": "images/s",
        "ms_per_step": wall / args.steps * 1e3, "serial_ms_per_step": wall_serial / args.steps * 1e3,
        "steps": args.steps, "warmup": args.warmup, "stages_ms": stages,
        "config": {"workload": "detect + Analyzer (merge, best-of-overlap, pixel lists), %d maps at %d, score_thr=median" % (B, S),
                   "score_thr": an.score_thr, "objects_per_image": n_obj / (B * args.steps),
                   "pixels_per_image": n_pix / (B * args.steps)},
        "e2e": {"h2d_bytes_per_step": B * S * S * 4,
                "d2h_bytes_per_step": B * D * (16 + 4 + 4) + B * 4 + int(n_pix / args.steps) * 8,
                "masks_bytes_kept_on_device_per_step": B * S * S * D},
    }))


if __name__ == "__main__":
    main()
