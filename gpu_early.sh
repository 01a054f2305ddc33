#!/bin/bash
# scratch driver for one gpurun call: analyzer GPU tests + analyzer bench + catalogue bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_analyzer.py -q -m gpu -x > gpurun_out/tests_an.log 2>&1; echo "tests exit $?" >> gpurun_out/tests_an.log; tail -25 gpurun_out/tests_an.log
timeout 600 python tools/analyze_bench.py > gpurun_out/analyze_bench.log 2> gpurun_out/analyze_bench.err; echo "abench exit $?"; tail -1 gpurun_out/analyze_bench.log | cut -c1-1800; tail -5 gpurun_out/analyze_bench.err
timeout 600 python tools/catalog_bench.py > gpurun_out/catalog_bench.log 2> gpurun_out/catalog_bench.err; echo "cbench exit $?"; tail -1 gpurun_out/catalog_bench.log | cut -c1-1500; tail -5 gpurun_out/catalog_bench.err
