#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/tests.log 2>&1; echo "tests exit $?" >> gpurun_out/tests.log; tail -5 gpurun_out/tests.log
timeout 600 python tools/layer_table.py 64 > gpurun_out/layer_table.txt 2>&1
grep "stem\|roialign\|proposal\|detection\|total" gpurun_out/layer_table.txt | cut -c1-80
BENCH_DEBUG=1 timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','e2e','gpu_launches')})
print(d['roofline']['frac']); print(d['stage_ms_per_step']); print({k:round(v['ms_per_step'],4) for k,v in d['kernel_families'].items()})
PY
tail -3 gpurun_out/bench.err
