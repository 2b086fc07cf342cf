#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_backward.py tests/test_gpu_network.py -x -q -m gpu 2>&1 | tail -6
timeout 180 python scripts/bench_pair.py 2>&1 | tail -20
