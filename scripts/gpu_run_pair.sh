#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_backward.py tests/test_gpu_network.py tests/test_gpu_training.py -x -q -m gpu 2>&1 | tail -12 > gpurun_out/pair_tests.txt
cat gpurun_out/pair_tests.txt
timeout 300 python bench.py --workload train_seg --height 320 --width 640 --batch 16 --steps 10 --warmup 3 > gpurun_out/train_seg.json 2> gpurun_out/train_seg.err; tail -c 400 gpurun_out/train_seg.json; tail -3 gpurun_out/train_seg.err
timeout 300 python bench.py --workload train_critic --height 320 --width 640 --batch 16 --steps 10 --warmup 3 > gpurun_out/train_critic.json 2> gpurun_out/train_critic.err; tail -c 400 gpurun_out/train_critic.json; tail -3 gpurun_out/train_critic.err
