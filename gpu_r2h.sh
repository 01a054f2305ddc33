#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_training.py -q -m gpu -x -k "graph or steps" 2>&1 | tail -30 > gpurun_out/train_tests.log; tail -25 gpurun_out/train_tests.log
MRCNN_B200_TRAIN_BACKEND=torch timeout 900 python -m pytest tests/test_gpu_training.py -q -m gpu -x -k "graph or steps" 2>&1 | tail -3
for be in tcgen05 torch; do
MRCNN_B200_TRAIN_BACKEND=$be timeout 600 python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/train_bench_$be.log 2> gpurun_out/train_bench_$be.err; echo "train bench $be exit $?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/train_bench_$be.log').read().strip().splitlines()[-1])
    print(round(d['value'],1),'img/s', round(d['ms_per_step'],2),'ms/step e2e',round(d['e2e']['ms_per_step'],2), d['phase_ms_per_step_eager'], d['config']['cuda_graph'], d['losses_last_step'])
except Exception as e:
    print('no json', e); print(open('gpurun_out/train_bench_$be.err').read()[-1500:])
PY
done
