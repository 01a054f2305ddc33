#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_training.py -q -m gpu -x 2>&1 | tail -40 > gpurun_out/train_tests.log; cat gpurun_out/train_tests.log | tail -40
