#!/bin/bash
# final pass: GPU tests, smoke, bench, analyzer / catalogue / tile timings, ncu list of the analyzer kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/tests.log 2>&1; echo "tests exit $?" >> gpurun_out/tests.log; grep -v "Invalid det bbox" gpurun_out/tests.log | tail -4
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; tail -1 gpurun_out/bench.log | cut -c1-260
timeout 600 python tools/analyze_bench.py > gpurun_out/analyze_bench.log 2> gpurun_out/analyze_bench.err; echo "abench exit $?"; tail -1 gpurun_out/analyze_bench.log | cut -c1-1500
timeout 600 python tools/catalog_bench.py > gpurun_out/catalog_bench.log 2> gpurun_out/catalog_bench.err; echo "cbench exit $?"; tail -1 gpurun_out/catalog_bench.log | cut -c1-1200
timeout 600 python tools/tile_bench.py > gpurun_out/tile_bench.log 2> gpurun_out/tile_bench.err; echo "tbench exit $?"; tail -1 gpurun_out/tile_bench.log | cut -c1-700
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"masks_pack|planes_|label_|labels_|pixel_lists" -c 14 --csv --log-file gpurun_out/analyzer_launches.csv python tools/analyze_bench.py --reps 1 --oracle-frames 0 > gpurun_out/ncu_an_light.log 2>&1; echo "ncu exit $?"
