#!/bin/bash
mkdir -p gpurun_out
export MRCNN_B200_AUTOTUNE_CACHE=$PWD/gpurun_out/autotune_cache_r2.txt
rm -f $MRCNN_B200_AUTOTUNE_CACHE
timeout 300 python tools/profile_run.py 64 2 > gpurun_out/plain_r2.log 2>&1 || { tail -5 gpurun_out/plain_r2.log; exit 1; }
tail -1 gpurun_out/plain_r2.log
timeout 900 ncu --set full --clock-control none -k regex:"conv_gemm_kernel" -s 164 -c 3 -f -o gpurun_out/prof_res4 python tools/profile_run.py 64 2 > gpurun_out/ncu_res4.log 2>&1; echo "res4 exit $?"
ncu -i gpurun_out/prof_res4.ncu-rep --page raw --csv > gpurun_out/prof_res4.raw.csv 2>/dev/null
ncu -i gpurun_out/prof_res4.ncu-rep --page details --csv > gpurun_out/prof_res4.details.csv 2>/dev/null
rm -f gpurun_out/prof_res4.ncu-rep
ls -la gpurun_out
