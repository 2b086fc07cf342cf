# final-build evidence: launch list of the inference step, --set full capture of the CTA-pair layers
set -x
mkdir -p gpurun_out
CMD2="python bench.py --legs none --no-logits-e2e --steps 1 --warmup 3 --no-cpu-baseline"
timeout 300 $CMD2 > gpurun_out/infer_plain.json 2> gpurun_out/infer_plain.err || exit 1
HN_PROFILE_RANGE=1 timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_infer.csv $CMD2 > gpurun_out/ncu_infer.log 2>&1
tail -2 gpurun_out/ncu_infer.log | cut -c1-200
timeout 120 python scripts/ncu_pair_layers.py > gpurun_out/pair_layers_plain.log 2>&1 || exit 1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_ -o gpurun_out/r2_full_pair -f python scripts/ncu_pair_layers.py > gpurun_out/ncu_full_pair.log 2>&1; tail -2 gpurun_out/ncu_full_pair.log | cut -c1-200
