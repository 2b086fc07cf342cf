#!/bin/bash
# CTA-pair convolution kernels: layer timings with / without pairs, role waits, then the kernel parity tests
mkdir -p gpurun_out
{
for env in "HN_NO_PAIR=1" "HN_PAIR=1" "HN_PAIR_MIN_KB=1"; do
  echo "=== $env"
  env $env timeout 180 python scripts/bench_pair.py 2>&1 | tail -18
done
env HN_PAIR=1 timeout 180 python scripts/profile_halo_pair.py 2>&1 | tail -6
} > gpurun_out/bench_pair.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_backward.py tests/test_gpu_network.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/pair_tests.txt
cat gpurun_out/bench_pair.txt gpurun_out/pair_tests.txt
