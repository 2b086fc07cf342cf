timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_training.py -m gpu -q 2>&1 | grep -E "^E  |passed|failed|^FAILED" | cut -c1-250 | head -30
HN_TIMELINE=1 python bench.py --workload train_seg --height 320 --width 640 --batch 16 --steps 5 --warmup 3 --layer-table gpurun_out/train_seg_calls4.json > gpurun_out/train_seg12.json 2> gpurun_out/train_seg12.err
python bench.py --workload train_seg --height 320 --width 640 --batch 16 --steps 5 --warmup 3 > gpurun_out/train_seg13.json 2> gpurun_out/train_seg13.err; tail -c 300 gpurun_out/train_seg13.json
