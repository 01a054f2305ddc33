#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_graph_layers.py tests/test_gpu_engine.py -q -m gpu -x > gpurun_out/tests.log 2>&1; echo "tests exit $?" >> gpurun_out/tests.log; tail -3 gpurun_out/tests.log
MRCNN_B200_PROPOSAL_CLOCKS=1 timeout 300 python tools/proposal_stats.py 2>&1 | grep "proposal phases" | tail -1
timeout 600 python tools/layer_table.py 64 > gpurun_out/layer_table.txt 2>&1
grep "proposal\|total" gpurun_out/layer_table.txt
