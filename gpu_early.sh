#!/bin/bash
# scratch: development GPU run
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -q -m gpu > gpurun_out/conv.log 2>&1; echo "conv exit $?" >> gpurun_out/conv.log
timeout 1200 python -m pytest tests/test_gpu_engine.py -q -m gpu -x > gpurun_out/engine.log 2>&1; echo "engine exit $?" >> gpurun_out/engine.log
tail -15 gpurun_out/conv.log; tail -60 gpurun_out/engine.log
