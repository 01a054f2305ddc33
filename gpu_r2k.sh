#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py -q -m gpu -x 2>&1 | tail -4
./gpu_multi2.sh 2
