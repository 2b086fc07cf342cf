# ncu --set full of the pair-form weight-gradient kernel inside one eager train_seg step
set -x
mkdir -p gpurun_out
CMD="python bench.py --workload train_seg --height 320 --width 640 --batch 16 --steps 1 --warmup 3 --cuda-graph off"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 || exit 1
export HN_PROFILE_RANGE=1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -c 12 -o gpurun_out/r2_full_wgrad_pair -f $CMD > gpurun_out/ncu_full_wgrad.log 2>&1; tail -2 gpurun_out/ncu_full_wgrad.log | cut -c1-200
