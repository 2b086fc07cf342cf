# Four-GPU check of the default bench line (what the driver's scaling run launches):  gpurun --gpus 4 --timeout 900 -- 'bash scripts/gpu_run_n4.sh'
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513"
timeout 500 $TR bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/n4_bench.json 2> gpurun_out/n4_bench.err; echo rc=$?; tail -c 1200 gpurun_out/n4_bench.json; grep -i "error\|Traceback" gpurun_out/n4_bench.err | head -5
