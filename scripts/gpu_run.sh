set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
python __graft_entry__.py --smoke 2>&1 | tail -3
python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; tail -c 400 gpurun_out/bench_final2.json
python bench.py --workload iou_eval > gpurun_out/iou_final.json 2> /dev/null
python bench.py --workload train_seg --height 320 --width 640 --batch 16 --steps 5 --warmup 3 > gpurun_out/train_seg_final.json 2> /dev/null
python bench.py --workload train_critic --height 320 --width 640 --batch 16 --steps 5 --warmup 3 > gpurun_out/train_critic_final.json 2> /dev/null
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/ref_final.json 2> /dev/null; tail -c 300 gpurun_out/ref_final.json
