#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/tile_bench.py > gpurun_out/tile_bench.log 2> gpurun_out/tile_bench.err; echo "tbench exit $?"; tail -1 gpurun_out/tile_bench.log | cut -c1-1500; grep -v "Invalid det bbox\|No object detected" gpurun_out/tile_bench.err | tail -8
