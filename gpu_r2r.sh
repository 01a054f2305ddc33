#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_graph_layers.py tests/test_gpu_engine.py -q -m gpu -x 2>&1 | tail -4
MRCNN_B200_PROPOSAL_CLOCKS=1 timeout 600 python bench.py --steps 4 --warmup 1 --no-cpu-baseline 2>&1 >/dev/null | grep "proposal phases" | tail -3
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],3), {k: round(v['ms_per_step'],4) for k,v in d['kernel_families'].items()})"
