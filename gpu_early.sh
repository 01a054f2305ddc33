#!/bin/bash
# scratch: development GPU run
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -q -m gpu > gpurun_out/conv.log 2>&1; echo "conv exit $?" >> gpurun_out/conv.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
tail -3 gpurun_out/conv.log; tail -5 gpurun_out/smoke.log; cat gpurun_out/bench.log; tail -20 gpurun_out/bench.err
