# Standard verification on a B200 box:  gpurun --timeout 1500 -- 'bash scripts/gpu_run.sh'
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/pytest.log 2>&1; tail -25 gpurun_out/pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -8 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
