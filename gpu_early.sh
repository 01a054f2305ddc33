#!/bin/bash
# scratch: development GPU run
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -q -m gpu -x > gpurun_out/tests.log 2>&1; echo "tests exit $?" >> gpurun_out/tests.log
tail -8 gpurun_out/tests.log
timeout 600 python tools/layer_table.py 64 > gpurun_out/layer_table.txt 2>&1
grep "roialign\|proposal\|detection\|total" gpurun_out/layer_table.txt
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','clocks','cpu_baseline')})
print(d['roofline']); print(d['stage_ms_per_step'])
PY
tail -3 gpurun_out/bench.err
