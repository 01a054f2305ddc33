#!/bin/bash
mkdir -p gpurun_out
run() { BENCH_SPARSE=0 timeout 600 python bench.py --steps 30 --warmup 6 --no-cpu-baseline 2>gpurun_out/h.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'e2e ms', round(d['e2e']['ms_per_step'],2), 'share', d['e2e'].get('dense_share_images'), 'packed', round(d['e2e_packed_masks']['value'],1))" || tail -5 gpurun_out/h.err; }
run "threads=16 default (off)"
MRCNN_B200_HOST_THREADS=4 MRCNN_B200_DENSE_SHARE=auto run "threads=4 auto"
MRCNN_B200_HOST_THREADS=2 MRCNN_B200_DENSE_SHARE=auto run "threads=2 auto"
MRCNN_B200_HOST_THREADS=1 MRCNN_B200_DENSE_SHARE=auto run "threads=1 auto"
timeout 900 python -m pytest tests/test_gpu_engine.py -q -m gpu -x -k "hybrid or expand or async or packed or end_to_end" 2>&1 | tail -3
