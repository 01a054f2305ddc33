#!/bin/bash
mkdir -p gpurun_out
run() { timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/w.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value'],1), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stage_ms_per_step'].items()})"; }
true
MRCNN_B200_AUTOTUNE=0 MRCNN_B200_OCC2=2 run "occ2-all-296"
MRCNN_B200_AUTOTUNE=0 MRCNN_B200_OCC2=2 MRCNN_B200_OCC2_GRID=1 run "occ2-all-148"
tail -5 gpurun_out/w.err
