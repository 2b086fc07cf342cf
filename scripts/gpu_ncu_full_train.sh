# ncu --set full captures of the backward-path kernels of one eager train_seg step:  gpurun --timeout 1500 -- 'bash scripts/gpu_ncu_full_train.sh'
set -x
mkdir -p gpurun_out
CMD="python bench.py --workload train_seg --height 320 --width 640 --batch 16 --steps 1 --warmup 3 --cuda-graph off"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 || exit 1
export HN_PROFILE_RANGE=1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -s 40 -c 4 -o gpurun_out/r2_full_wgrad -f $CMD > gpurun_out/ncu_full_wgrad.log 2>&1; tail -2 gpurun_out/ncu_full_wgrad.log | cut -c1-200
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:bn_ -s 200 -c 8 -o gpurun_out/r2_full_bn -f $CMD > gpurun_out/ncu_full_bn.log 2>&1; tail -2 gpurun_out/ncu_full_bn.log | cut -c1-200
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 330 -c 6 -o gpurun_out/r2_full_conv_bwd -f $CMD > gpurun_out/ncu_full_conv.log 2>&1; tail -2 gpurun_out/ncu_full_conv.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
