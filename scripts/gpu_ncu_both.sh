set -x
mkdir -p gpurun_out
# launch lists (gpu__time_duration) of the train_seg step and of the inference step, after each command has run clean without ncu
CMD="python bench.py --workload train_seg --height 320 --width 640 --batch 16 --steps 1 --warmup 3 --cuda-graph off"
timeout 300 $CMD > gpurun_out/train_plain.json 2> gpurun_out/train_plain.err || exit 1
HN_PROFILE_RANGE=1 timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv $CMD > gpurun_out/ncu_train.log 2>&1
tail -2 gpurun_out/ncu_train.log | cut -c1-200
CMD2="python bench.py --legs none --no-logits-e2e --steps 1 --warmup 3 --no-cpu-baseline"
timeout 300 $CMD2 > gpurun_out/infer_plain.json 2> gpurun_out/infer_plain.err || exit 1
HN_PROFILE_RANGE=1 timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_infer.csv $CMD2 > gpurun_out/ncu_infer.log 2>&1
tail -2 gpurun_out/ncu_infer.log | cut -c1-200
