#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_conv.py -q -m gpu -x 2>&1 | tail -12
timeout 600 python -m pytest tests/test_gpu_engine.py -q -m gpu -x -k "error_conv" 2>&1 | tail -3
timeout 300 python bench.py --mode train --steps 20 --warmup 3 > gpurun_out/train_bench.log 2> gpurun_out/train_bench.err; echo "train bench exit $?"; python - <<PY
import json
d=json.loads(open('gpurun_out/train_bench.log').read().strip().splitlines()[-1])
print(round(d['value'],1),'img/s', round(d['ms_per_step'],2),'ms/step e2e',round(d['e2e']['ms_per_step'],2), d['phase_ms_per_step_eager'], d['config']['cuda_graph'])
PY
timeout 300 python tools/train_profile.py gpurun_out/train_profile.json 2>&1 | tail -24
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('detect', round(d['value'],1), round(d['ms_per_step'],3), round(d['roofline']['frac'],3))"
