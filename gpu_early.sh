#!/bin/bash
# scratch: development GPU run
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -q -m gpu -x > gpurun_out/conv.log 2>&1; echo "conv exit $?" >> gpurun_out/conv.log
tail -25 gpurun_out/conv.log
if grep -q "conv exit 0" gpurun_out/conv.log; then
timeout 900 python -m pytest tests/test_gpu_engine.py -q -m gpu -x > gpurun_out/tests.log 2>&1; echo "tests exit $?" >> gpurun_out/tests.log
timeout 600 python tools/layer_table.py 64 > gpurun_out/layer_table.txt 2>&1; echo "table exit $?" >> gpurun_out/layer_table.txt
tail -5 gpurun_out/tests.log; grep -v "res4[b-v]" gpurun_out/layer_table.txt
fi
