for i in 1 2 3; do timeout 600 python -m pytest tests/test_gpu_backward.py tests/test_gpu_kernels.py -m gpu -q 2>&1 | grep -E "^E  |^tests.*Error|passed|failed|^FAILED" | cut -c1-250 | head -30; done
