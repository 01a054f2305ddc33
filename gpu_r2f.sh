#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_training.py -q -m gpu -x 2>&1 | tail -40 > gpurun_out/train_tests.log; tail -30 gpurun_out/train_tests.log
timeout 600 python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/train_bench.log 2> gpurun_out/train_bench.err; echo "train bench exit $?"; tail -c 2500 gpurun_out/train_bench.log; tail -5 gpurun_out/train_bench.err
