set -x
mkdir -p gpurun_out
timeout 60 scripts/_bin/probe_rate2 > gpurun_out/probe_rate2.txt 2>&1; echo rc=$?; cat gpurun_out/probe_rate2.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -8 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 1500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
