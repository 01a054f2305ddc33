#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/tests.log 2>&1; echo "tests exit $?" >> gpurun_out/tests.log; tail -4 gpurun_out/tests.log
timeout 600 python tools/layer_table.py 64 > gpurun_out/layer_table.txt 2>&1
grep "total" gpurun_out/layer_table.txt
for pdl in 1 0; do
MRCNN_B200_PDL=$pdl timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pdl$pdl.log 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_pdl$pdl.log').read().strip().splitlines()[-1])
print("PDL=$pdl", {k:d.get(k) for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'])
print(d['stage_ms_per_step'])
PY
done
timeout 300 python tools/two_stream_probe.py 2>&1 | tail -3
