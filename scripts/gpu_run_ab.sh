#!/bin/bash
# A/B of one environment knob on the headline inference step and the train_seg step
mkdir -p gpurun_out
for tag in a b; do
  if [ $tag = a ]; then export HN_PAIR_MAX_BN=128; else export HN_PAIR_MAX_BN=256; fi
  timeout 300 python bench.py --legs none --no-logits-e2e > gpurun_out/ab_infer_$tag.json 2> gpurun_out/ab_infer_$tag.err
  timeout 300 python bench.py --workload train_seg --height 320 --width 640 --batch 16 --steps 10 --warmup 3 > gpurun_out/ab_train_$tag.json 2> gpurun_out/ab_train_$tag.err
  python - <<PY
import json
d=json.load(open('gpurun_out/ab_infer_$tag.json')); print('$tag infer', d['ms_per_step'], d['value'], d['roofline']['achieved'], d['clocks'])
d=json.load(open('gpurun_out/ab_train_$tag.json')); print('$tag train', d['ms_per_step'], d['value'], d['clocks'])
PY
done
