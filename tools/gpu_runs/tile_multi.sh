#!/bin/bash
mkdir -p gpurun_out
timeout 600 tools/tile_multi_check.sh 2>&1 | tail -5; echo "exit $?"
tail -3 gpurun_out/tile_multi.log | cut -c1-300
