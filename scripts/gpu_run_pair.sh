#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_bf16_training.py tests/test_gpu_network.py -x -q -m gpu 2>&1 | tail -12 > gpurun_out/pair_tests.txt
cat gpurun_out/pair_tests.txt
timeout 300 python bench.py --workload train_seg --height 320 --width 640 --batch 16 --steps 10 --warmup 3 > gpurun_out/train_seg.json 2> gpurun_out/train_seg.err; tail -c 300 gpurun_out/train_seg.json; tail -3 gpurun_out/train_seg.err
timeout 300 python bench.py --workload train_critic --height 320 --width 640 --batch 16 --steps 10 --warmup 3 > gpurun_out/train_critic.json 2> gpurun_out/train_critic.err; tail -c 300 gpurun_out/train_critic.json; tail -3 gpurun_out/train_critic.err
python - <<PY
import json
for f in ('train_seg','train_critic'):
    d=json.load(open(f'gpurun_out/{f}.json')); print(f, d['ms_per_step'], d['value'], d['clocks']['sm_mhz'], d.get('gpu_launches'))
PY
