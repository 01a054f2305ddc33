#!/bin/bash
# Round-1 measurement pass: bench (plain), launch list of one step, light metrics of every GEMM launch of one step,
# full ncu captures of selected GEMMs and of the non-GEMM kernels.  The autotune cache makes every process build
# the same launch plan.  Output kept under 64 MiB (gpurun's copy-back limit).
mkdir -p gpurun_out
export MRCNN_B200_AUTOTUNE_CACHE=$PWD/gpurun_out/autotune_cache.txt
rm -f $MRCNN_B200_AUTOTUNE_CACHE
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err || { tail -5 gpurun_out/bench.err; exit 1; }
tail -1 gpurun_out/bench.log | cut -c1-200
timeout 300 python tools/profile_run.py 64 2 > gpurun_out/plain.log 2>&1 || { tail -5 gpurun_out/plain.log; exit 1; }
NLAUNCH=$(python -c "print(open('gpurun_out/plain.log').read().split('launches_per_step=')[1].split()[0])")
echo "launches per step: $NLAUNCH"
# (1) every launch of the second step with its device time
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s $NLAUNCH -c $NLAUNCH --csv --log-file gpurun_out/launches.csv python tools/profile_run.py 64 2 > gpurun_out/ncu_list.log 2>&1
# (2) DRAM traffic + tensor-pipe activity of all 130 GEMM launches of the second step (few passes)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:conv_gemm_kernel -s 130 -c 130 --csv --log-file gpurun_out/gemm_all_light.csv python tools/profile_run.py 64 2 > gpurun_out/ncu_gemm_light.log 2>&1
# (3) full sets: class + mask head GEMMs (fc1, fc2, head, mask conv1-4, fused deconv+logits) and one res4 bottleneck
timeout 900 ncu --set full --clock-control none -k regex:conv_gemm_kernel -s 252 -c 8 -f -o gpurun_out/prof_gemm_heads python tools/profile_run.py 64 2 > gpurun_out/ncu_gemm_heads.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:conv_gemm_kernel -s 158 -c 3 -f -o gpurun_out/prof_gemm_res4b python tools/profile_run.py 64 2 > gpurun_out/ncu_gemm_res4b.log 2>&1
# (4) the non-GEMM kernels of the second step
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"roialign_kernel|proposal_kernel|detection_kernel|stem_im2col|zscale|stretch|resize_pad|unmold|maxpool" -s 13 -c 13 -f -o gpurun_out/prof_misc3 python tools/profile_run.py 64 2 > gpurun_out/ncu_misc3.log 2>&1
for r in prof_gemm_heads prof_gemm_res4b prof_misc3; do ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null; done
du -sh gpurun_out; ls -la gpurun_out
