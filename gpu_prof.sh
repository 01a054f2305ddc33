#!/bin/bash
# Round-1 measurement pass: bench (plain), launch list of one step, full ncu capture of every GEMM launch of one step
# and of the non-GEMM hot kernels.  The autotune cache makes every process build the same launch plan.
mkdir -p gpurun_out
export MRCNN_B200_AUTOTUNE_CACHE=$PWD/gpurun_out/autotune_cache.txt
rm -f $MRCNN_B200_AUTOTUNE_CACHE
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err || { tail -5 gpurun_out/bench.err; exit 1; }
tail -1 gpurun_out/bench.log | cut -c1-300
wc -l $MRCNN_B200_AUTOTUNE_CACHE
timeout 300 python tools/profile_run.py 64 2 > gpurun_out/plain.log 2>&1 || { tail -5 gpurun_out/plain.log; exit 1; }
# (1) every launch of the second step with its device time
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 156 -c 170 --csv --log-file gpurun_out/launches.csv python tools/profile_run.py 64 2 > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log
# (2) all GEMM launches of the second step, full set
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s 130 -c 130 -f -o gpurun_out/prof_gemm_all python tools/profile_run.py 64 2 > gpurun_out/ncu_gemm.log 2>&1
tail -2 gpurun_out/ncu_gemm.log
# (3) the non-GEMM kernels of the second step
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"roialign_kernel|proposal_kernel|detection_kernel|stem_im2col|zscale|stretch|resize_pad|unmold|maxpool" -s 13 -c 13 -f -o gpurun_out/prof_misc3 python tools/profile_run.py 64 2 > gpurun_out/ncu_misc3.log 2>&1
tail -2 gpurun_out/ncu_misc3.log
ls -la gpurun_out/*.ncu-rep
