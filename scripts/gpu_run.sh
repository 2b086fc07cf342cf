timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_training.py tests/test_gpu_network.py -m gpu -x -q 2>&1 | grep -E "^E  |passed|failed|^FAILED|Error" | cut -c1-300 | head -20
python bench.py --workload train_seg --height 320 --width 640 --batch 16 --steps 5 --warmup 3 > gpurun_out/train_seg17.json 2> gpurun_out/train_seg17.err; tail -c 200 gpurun_out/train_seg17.json
