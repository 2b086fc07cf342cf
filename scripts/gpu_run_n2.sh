# Two-GPU check of the gradient exchange:  gpurun --gpus 2 --timeout 900 -- 'bash scripts/gpu_run_n2.sh'
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err; echo rc=$?; tail -c 2500 gpurun_out/n2_bench.json; grep "rank0\]:" gpurun_out/n2_bench.err | tail -6
