#!/bin/bash
# scratch driver for one gpurun call: tile-driver GPU tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sfinder.py -q -m gpu -x > gpurun_out/tests_sf.log 2>&1; echo "tests exit $?" >> gpurun_out/tests_sf.log; grep -v "Invalid det bbox" gpurun_out/tests_sf.log | tail -40
