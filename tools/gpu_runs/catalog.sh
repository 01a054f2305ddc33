#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_analyzer.py tests/test_gpu_sfinder.py -q -m gpu -x 2>&1 | tail -5
timeout 600 python tools/catalog_bench.py --steps 20 --warmup 4 > gpurun_out/catalog_bench.json 2> gpurun_out/catalog_bench.err; echo "exit $?"; tail -3 gpurun_out/catalog_bench.err; cat gpurun_out/catalog_bench.json
