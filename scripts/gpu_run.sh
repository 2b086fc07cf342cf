set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/pytest.log 2>&1; tail -8 gpurun_out/pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -8 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 1500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
HN_NO_PAIR=1 timeout 300 python bench.py --legs none > gpurun_out/bench_single_cta.json 2> gpurun_out/bench_single_cta.err; tail -c 600 gpurun_out/bench_single_cta.json
