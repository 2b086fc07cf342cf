#!/bin/bash
# A/B of one environment knob on the headline inference step (and optionally the train_seg step):  VAR=HN_PAIR_MIN_KB VALS="9 8 4" bash scripts/gpu_run_ab.sh
mkdir -p gpurun_out
VAR=${VAR:-HN_PAIR_MIN_KB}
for v in ${VALS:-9 8 4}; do
  if [ "$v" = unset ]; then unset $VAR; else export $VAR=$v; fi          # (the library's switches test presence, not value)
  timeout 300 python bench.py --legs none --no-logits-e2e --no-cpu-baseline > gpurun_out/ab_infer_$v.json 2> gpurun_out/ab_infer_$v.err
  if [ -n "$TRAIN" ]; then timeout 300 python bench.py --workload train_seg --height 320 --width 640 --batch 16 --steps 10 --warmup 3 > gpurun_out/ab_train_$v.json 2> gpurun_out/ab_train_$v.err; fi
  python - <<PY
import json, os
d=json.load(open('gpurun_out/ab_infer_$v.json')); print('$VAR=$v infer', round(d['ms_per_step'],2), round(d['value'],1), round(d['roofline']['achieved']), d['clocks']['sm_mhz'], d['clocks'].get('power_w_max'))
if os.environ.get('TRAIN'):
    d=json.load(open('gpurun_out/ab_train_$v.json')); print('$VAR=$v train', round(d['ms_per_step'],2), round(d['value'],1), d['clocks']['sm_mhz'])
PY
done
