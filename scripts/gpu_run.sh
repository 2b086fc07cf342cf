timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k confusion 2>&1 | tail -3
python bench.py --workload iou_eval --no-cpu-baseline > gpurun_out/iou4.json 2> gpurun_out/iou4.err; tail -c 400 gpurun_out/iou4.json
