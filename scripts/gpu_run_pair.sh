#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_backward.py tests/test_gpu_network.py -x -q -m gpu 2>&1 | tail -12 > gpurun_out/pair_tests.txt
cat gpurun_out/pair_tests.txt
