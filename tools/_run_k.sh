timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_md.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --workload spce --lattice 32 --steps 50 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/bench_spce32_k4.json 2> gpurun_out/bench_spce32_k4.err; tail -3 gpurun_out/bench_spce32_k4.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_spce32_k4.json').read().splitlines()[-1]); print(d['value'], d['ms_per_step'], {k:v.get('ms_per_step', v.get('rho_plus_force_ms', v.get('avg_launch_ms'))) for k,v in d['roofline_extra'].items()})"
