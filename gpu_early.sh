#!/bin/bash
# scratch: first GPU run (development)
export MRCNN_B200_PARTIAL=1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_conv.py -x -q -m gpu > gpurun_out/conv.log 2>&1; echo "conv exit $?" >> gpurun_out/conv.log
timeout 600 python -m pytest tests/test_gpu_graph_layers.py -q -m gpu > gpurun_out/layers.log 2>&1; echo "layers exit $?" >> gpurun_out/layers.log
tail -30 gpurun_out/conv.log; tail -40 gpurun_out/layers.log
