#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_conv.py -q -m gpu -x > gpurun_out/tests.log 2>&1; echo "tests exit $?" >> gpurun_out/tests.log; tail -3 gpurun_out/tests.log
timeout 300 python tools/layer_table.py 64 > gpurun_out/layer_table.txt 2>&1
grep "mrcnn_mask (\|total" gpurun_out/layer_table.txt | cut -c1-90
