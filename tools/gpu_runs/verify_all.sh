#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu -x --durations=12 2>&1 | tail -30 ) 2>&1
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 ) 2>&1
( time timeout 600 python bench.py > gpurun_out/bench_s.log 2> gpurun_out/bench_s.err ) 2>&1 | tail -4; echo "bench exit $?"; tail -c 6000 gpurun_out/bench_s.log
