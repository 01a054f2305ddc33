#!/bin/bash
# Round-2 measurement pass (one GPU): launch list of one detect step, full ncu captures of the non-GEMM kernels that changed
# (ProposalLayer, row-per-warp ROIAlign, bit-plane unmold) and of the train-mode kernels (weight-gradient GEMM, backward
# prologue, ROIAlign backward, DetectionTargetLayer, optimiser).
mkdir -p gpurun_out
export MRCNN_B200_AUTOTUNE_CACHE=$PWD/gpurun_out/autotune_cache_r2.txt
rm -f $MRCNN_B200_AUTOTUNE_CACHE
timeout 300 python tools/profile_run.py 64 2 > gpurun_out/plain_r2.log 2>&1 || { tail -5 gpurun_out/plain_r2.log; exit 1; }
NLAUNCH=$(python -c "print(open('gpurun_out/plain_r2.log').read().split('launches_per_step=')[1].split()[0])")
echo "launches per step: $NLAUNCH"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s $NLAUNCH -c $NLAUNCH --csv --log-file gpurun_out/r02_launches.csv python tools/profile_run.py 64 2 > gpurun_out/ncu_list_r2.log 2>&1; echo "list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"roialign_rows_kernel|proposal_kernel|detection_kernel|unmold_paint_bits|unmold_boxes|roi_levels|mask_tile_flags|mask_zero_padded" -s 10 -c 10 -f -o gpurun_out/prof_r2_misc python tools/profile_run.py 64 2 > gpurun_out/ncu_r2_misc.log 2>&1; echo "misc exit $?"
NO_TORCH_PROF=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"wgrad_kernel|conv_backward_prep|roialign_backward|detection_targets|sgd_norm|sgd_apply" -s 745 -c 28 -f -o gpurun_out/prof_r2_train python tools/train_profile.py > gpurun_out/ncu_r2_train.log 2>&1; echo "train exit $?"
for r in prof_r2_misc prof_r2_train; do ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null; rm -f gpurun_out/$r.ncu-rep; done   # copy-back limit: 64 MiB
ls -la gpurun_out | grep r2
