#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -q -m gpu -x 2>&1 | tail -5
export MRCNN_B200_AUTOTUNE_CACHE=$PWD/gpurun_out/autotune_cache_occ2.txt
rm -f $MRCNN_B200_AUTOTUNE_CACHE
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_v.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],3), d['roofline']['frac'], d['stage_ms_per_step'])"
tail -3 gpurun_out/bench_v.err
awk '{print $3}' $MRCNN_B200_AUTOTUNE_CACHE | sort | uniq -c
MRCNN_B200_OCC2=0 MRCNN_B200_AUTOTUNE_CACHE= timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('no occ2:', round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],3), d['roofline']['frac'], d['stage_ms_per_step'])"
