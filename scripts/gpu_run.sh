timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --no-cpu-baseline > gpurun_out/bench20.json 2> gpurun_out/bench20.err; tail -c 300 gpurun_out/bench20.json; tail -5 gpurun_out/bench20.err
