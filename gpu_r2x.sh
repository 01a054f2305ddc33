#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_conv.py tests/test_gpu_graph_layers.py -q -m gpu -x 2>&1 | tail -8
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_x.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],3), d['roofline']['frac'], d['stage_ms_per_step']); print(d.get('sparse_detections')); print(d['gpu_launches'])"
tail -3 gpurun_out/bench_x.err
