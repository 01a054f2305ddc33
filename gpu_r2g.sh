#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_training.py -q -m gpu -k "wgrad" 2>&1 | tail -30 > gpurun_out/wgrad_tests.log; tail -30 gpurun_out/wgrad_tests.log
timeout 1500 python -m pytest tests/test_gpu_training.py -q -m gpu -x -k "not wgrad" 2>&1 | tail -30 > gpurun_out/train_tests.log; tail -12 gpurun_out/train_tests.log
