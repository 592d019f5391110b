set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --steps 300 --warmup 20 > gpurun_out/bench_lj_v4.json 2> gpurun_out/bench_lj_v4.err; tail -c 3500 gpurun_out/bench_lj_v4.json; tail -5 gpurun_out/bench_lj_v4.err
