set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 300 --warmup 20 > gpurun_out/bench_lj_base.json 2> gpurun_out/bench_lj_base.err; tail -c 3000 gpurun_out/bench_lj_base.json
timeout 600 python bench.py --workload spce --lattice 32 --steps 50 --warmup 5 > gpurun_out/bench_spce32_base.json 2> gpurun_out/bench_spce32_base.err; tail -c 3000 gpurun_out/bench_spce32_base.json; tail -5 gpurun_out/bench_spce32_base.err
