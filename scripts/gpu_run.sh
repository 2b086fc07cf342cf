# Standard verification on a B200 box:  gpurun --timeout 1200 -- 'bash scripts/gpu_run.sh'
set -x
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 400 gpurun_out/bench.json
timeout 300 python bench.py --workload iou_eval > gpurun_out/iou_eval.json 2> /dev/null
timeout 300 python bench.py --workload train_seg --height 320 --width 640 --batch 16 --steps 5 --warmup 3 > gpurun_out/train_seg.json 2> /dev/null
timeout 300 python bench.py --workload train_critic --height 320 --width 640 --batch 16 --steps 5 --warmup 3 > gpurun_out/train_critic.json 2> /dev/null
