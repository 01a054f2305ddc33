#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sfinder.py -q -m gpu -x > gpurun_out/tests_sf.log 2>&1; echo "tests exit $?" >> gpurun_out/tests_sf.log; grep -v "Invalid det bbox" gpurun_out/tests_sf.log | tail -4
timeout 600 python tools/tile_bench.py > gpurun_out/tile_bench.log 2> gpurun_out/tile_bench.err; echo "tbench exit $?"; tail -1 gpurun_out/tile_bench.log | cut -c1-600
