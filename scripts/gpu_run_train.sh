set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/pytest.log 2>&1; tail -15 gpurun_out/pytest.log
timeout 300 python bench.py --workload train_seg --height 320 --width 640 --batch 16 --steps 10 --warmup 3 > gpurun_out/train_seg.json 2> gpurun_out/train_seg.err; tail -c 300 gpurun_out/train_seg.json; tail -3 gpurun_out/train_seg.err
timeout 300 python bench.py --workload train_critic --height 320 --width 640 --batch 16 --steps 10 --warmup 3 > gpurun_out/train_critic.json 2> gpurun_out/train_critic.err; tail -c 300 gpurun_out/train_critic.json; tail -3 gpurun_out/train_critic.err
HN_NO_PAIR_FORWARD=1 timeout 300 python bench.py --workload train_seg --height 320 --width 640 --batch 16 --steps 10 --warmup 3 > gpurun_out/train_seg_nopair.json 2> gpurun_out/train_seg_nopair.err; tail -c 300 gpurun_out/train_seg_nopair.json
