#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/profile_run.py 64 2 > gpurun_out/plain.log 2>&1 || { tail -5 gpurun_out/plain.log; exit 1; }
# non-GEMM kernels of the second step (skip the first step's launches): roialign x2, proposal, detection
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"roialign_kernel|proposal_kernel|detection_kernel" --launch-skip 4 -c 4 -f -o gpurun_out/prof_misc2 python tools/profile_run.py 64 2 > gpurun_out/ncu_misc2.log 2>&1
tail -3 gpurun_out/ncu_misc2.log
