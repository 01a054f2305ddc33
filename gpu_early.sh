#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_engine.py -q -m gpu -x > gpurun_out/tests.log 2>&1; echo "tests exit $?" >> gpurun_out/tests.log; tail -4 gpurun_out/tests.log
timeout 300 python tools/layer_table.py 64 > gpurun_out/layer_table.txt 2>&1
grep "res2b\|res3b\|res4b\|res5b\|C[2345] \|total" gpurun_out/layer_table.txt | cut -c1-90
grep "res[2345]\|C[2345] " gpurun_out/layer_table.txt | awk '{s+=$3} END {print "sum of res layers:", s}'
