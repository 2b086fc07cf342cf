timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_network.py -m gpu -q 2>&1 | grep -E "^E  |passed|failed|^FAILED" | cut -c1-250 | head -30
python bench.py --layer-table gpurun_out/layers_r1q.json --no-cpu-baseline > gpurun_out/bench19.json 2> gpurun_out/bench19.err; tail -c 500 gpurun_out/bench19.json; tail -3 gpurun_out/bench19.err
