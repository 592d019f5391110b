timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_default5.json 2> gpurun_out/bench_default5.err; tail -3 gpurun_out/bench_default5.err
