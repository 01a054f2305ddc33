#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_analyzer.py -q -m gpu -x -k "pack or goldens or predict_maps_device" > gpurun_out/tests_an.log 2>&1; echo "tests exit $?" >> gpurun_out/tests_an.log; grep -v "Invalid det" gpurun_out/tests_an.log | tail -5
timeout 600 python tools/analyze_bench.py --oracle-frames 0 > gpurun_out/analyze_bench2.log 2> gpurun_out/analyze_bench2.err; echo "abench exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/analyze_bench2.log').read().strip().splitlines()[-1]); print(d['kernels'], d['ms_per_batch'])"
