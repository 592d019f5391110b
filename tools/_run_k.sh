timeout 300 python bench.py --workload spce --lattice 32 --steps 50 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/bench_spce32_k3.json 2> gpurun_out/bench_spce32_k3.err; tail -3 gpurun_out/bench_spce32_k3.err; python -c "
import json; d=json.load(open('gpurun_out/bench_spce32_k3.json')); print(d['value'], d['ms_per_step'], d['roofline'], d['roofline_extra'])"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file gpurun_out/r1p_spce_launches.csv python tools/profile_step.py --workload spce --lattice 32 --steps 4 2>&1 | tail -1
