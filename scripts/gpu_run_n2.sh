# Two-GPU check of the gradient exchange:  gpurun --gpus 2 --timeout 900 -- 'bash scripts/gpu_run_n2.sh'
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 200 $TR bench.py --gpus 2 --workload train_seg --height 320 --width 640 --batch 16 --steps 5 --warmup 3 --cuda-graph off > gpurun_out/n2_train_seg_eager.json 2> gpurun_out/n2_train_seg_eager.err; rc=$?; echo rc=$rc; tail -c 1500 gpurun_out/n2_train_seg_eager.json; grep -v "^$" gpurun_out/n2_train_seg_eager.err | grep "rank0\]:" | tail -6
if [ $rc -ne 0 ]; then exit 0; fi
timeout 240 $TR bench.py --gpus 2 --workload train_seg --height 320 --width 640 --batch 16 --steps 5 --warmup 3 > gpurun_out/n2_train_seg.json 2> gpurun_out/n2_train_seg.err; rc=$?; echo rc=$rc; tail -c 1500 gpurun_out/n2_train_seg.json; grep "rank0\]:" gpurun_out/n2_train_seg.err | tail -6
if [ $rc -ne 0 ]; then exit 0; fi
timeout 400 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err; echo rc=$?; tail -c 2500 gpurun_out/n2_bench.json; grep "rank0\]:" gpurun_out/n2_bench.err | tail -6
