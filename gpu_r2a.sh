#!/bin/bash
# round-2 call A: box info, host-expansion bench, GPU tests, smoke, bench (N=1)
mkdir -p gpurun_out
{ nproc; lscpu | grep -E "Model name|Socket|NUMA|Thread|Core|Flags" | sed 's/\(Flags:\).*\(avx512bw\).*/\1 ... \2 .../'; free -g | head -2; } > gpurun_out/box.txt 2>&1
timeout 120 python tools/host_expand_bench.py gpurun_out/host_expand.json > gpurun_out/host_expand.log 2>&1; tail -7 gpurun_out/host_expand.log | head -6
MRCNN_PARITY_JSON=gpurun_out/e2e_parity_gpu.json timeout 1500 python -m pytest tests -q -m gpu -x -s > gpurun_out/tests.log 2>&1; echo "tests exit $?" >> gpurun_out/tests.log; grep -v "Invalid det bbox" gpurun_out/tests.log | tail -5
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; tail -1 gpurun_out/bench.log | cut -c1-1500
