#!/bin/bash
# scratch: development GPU run
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -q -m gpu -x > gpurun_out/tests.log 2>&1; echo "tests exit $?" >> gpurun_out/tests.log
tail -15 gpurun_out/tests.log
timeout 600 python tools/layer_table.py 64 > gpurun_out/layer_table.txt 2>&1
grep "roialign\|proposal\|total" gpurun_out/layer_table.txt
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_short.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu.log; wc -l gpurun_out/launches.csv
