timeout 600 python -m pytest tests/test_gpu_network.py -m gpu -q 2>&1 | grep -E "^E  |passed|failed|^FAILED|Error" | cut -c1-300 | head -20
rm -f gpurun_out/sweep2.jsonl
for B in 1 2 4 8; do python bench.py --batch $B --height 320 --width 640 --steps 20 --no-cpu-baseline 2>/dev/null | grep '^{' >> gpurun_out/sweep2.jsonl; done
for B in 1 2; do python bench.py --batch $B --height 650 --width 1920 --steps 20 --no-cpu-baseline --cuda-graph on 2>/dev/null | grep '^{' >> gpurun_out/sweep2.jsonl; done
python bench.py --steps 10 --no-cpu-baseline --cuda-graph on 2>/dev/null | grep '^{' >> gpurun_out/sweep2.jsonl
wc -l gpurun_out/sweep2.jsonl
