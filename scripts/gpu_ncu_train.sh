set -x
mkdir -p gpurun_out
CMD="python bench.py --workload train_seg --height 320 --width 640 --batch 16 --steps 1 --warmup 3 --cuda-graph off"
timeout 300 $CMD > gpurun_out/train_plain.json 2> gpurun_out/train_plain.err || exit 1
HN_PROFILE_RANGE=1 timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv $CMD > gpurun_out/ncu_train.log 2>&1
tail -2 gpurun_out/ncu_train.log | cut -c1-200
