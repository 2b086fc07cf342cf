timeout 300 python -m pytest tests/test_gpu_network.py -m gpu -q -k cuda_graph 2>&1 | grep -E "^E  |passed|failed|Error" | cut -c1-300 | head
timeout 200 python bench.py --batch 1 --height 320 --width 640 --steps 20 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['cuda_graph'], round(d['ms_per_step'],3), round(d['e2e']['value'],1))"
