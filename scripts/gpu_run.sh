set -x
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 400 python bench.py > gpurun_out/bench_final3.json 2> gpurun_out/bench_final3.err; wc -l gpurun_out/bench_final3.json
timeout 300 python bench.py --workload train_seg --height 320 --width 640 --batch 16 --steps 5 --warmup 3 > gpurun_out/train_seg_final2.json 2> /dev/null
timeout 300 python bench.py --workload train_critic --height 320 --width 640 --batch 16 --steps 5 --warmup 3 > gpurun_out/train_critic_final2.json 2> /dev/null
