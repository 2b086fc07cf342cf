timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_training.py -m gpu -q 2>&1 | grep -E "^E  |passed|failed|^FAILED" | cut -c1-250 | head -30
python bench.py --workload train_seg --height 320 --width 640 --batch 16 --steps 5 --warmup 3 > gpurun_out/train_seg11.json 2> gpurun_out/train_seg11.err; tail -c 400 gpurun_out/train_seg11.json
