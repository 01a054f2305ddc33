#!/bin/bash
# scratch driver for one gpurun call: analyzer GPU tests + CLI test + analyzer bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_analyzer.py tests/test_gpu_engine.py -q -m gpu -x -k "analyzer or pack or pair or labels or helper or predict_maps or cli" > gpurun_out/tests_an.log 2>&1; echo "tests exit $?" >> gpurun_out/tests_an.log; tail -25 gpurun_out/tests_an.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python tools/analyze_bench.py > gpurun_out/analyze_bench.log 2> gpurun_out/analyze_bench.err; echo "abench exit $?"; tail -1 gpurun_out/analyze_bench.log | cut -c1-1800; tail -5 gpurun_out/analyze_bench.err
