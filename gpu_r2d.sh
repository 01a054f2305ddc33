#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_graph_layers.py -q -m gpu -x -k proposal > gpurun_out/prop_tests.log 2>&1; echo "proposal tests exit $?"; tail -3 gpurun_out/prop_tests.log
MRCNN_B200_PROPOSAL_CLOCKS=1 timeout 600 python bench.py --steps 4 --warmup 1 --no-cpu-baseline 2>&1 >/dev/null | grep "proposal phases" | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3), {k: round(v['ms_per_step'],4) for k,v in d['kernel_families'].items()})"
