timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_network.py -m gpu -x -q 2>&1 | tail -4
python bench.py --layer-table gpurun_out/layers_r1p.json --no-cpu-baseline > gpurun_out/bench18.json 2> gpurun_out/bench18.err; tail -c 500 gpurun_out/bench18.json
HN_UP2_NO_WALK=1 python bench.py --no-cpu-baseline > gpurun_out/bench18_nowalk.json 2> gpurun_out/bench18_nowalk.err; tail -c 500 gpurun_out/bench18_nowalk.json
