timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/bench_default4.json 2> gpurun_out/bench_default4.err; tail -3 gpurun_out/bench_default4.err
