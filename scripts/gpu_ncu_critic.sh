set -x
mkdir -p gpurun_out
CMD="python bench.py --workload train_critic --height 320 --width 640 --batch 16 --steps 1 --warmup 3 --cuda-graph off"
timeout 300 $CMD > gpurun_out/critic_plain.json 2> gpurun_out/critic_plain.err || exit 1
HN_PROFILE_RANGE=1 timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_critic.csv $CMD > gpurun_out/ncu_critic.log 2>&1
tail -2 gpurun_out/ncu_critic.log | cut -c1-200
