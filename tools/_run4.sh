set -x
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lj_box or staged_lj_kernel_vs" 2>&1 | tail -5 || exit 1
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_md.py -m gpu -x -q -k "lj or staged or large" 2>&1 | tail -5
timeout 200 python bench.py --steps 300 --warmup 20 --no-e2e --no-cpu-baseline > gpurun_out/bench_lj_v9.json 2> gpurun_out/bench_lj_v9.err; python -c "
import json; d=json.load(open('gpurun_out/bench_lj_v9.json')); print(d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['roofline_extra']['neighbor_list'])"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:lj_force_kernel -s 2 -c 1 -o gpurun_out/r1m_lj_force python tools/profile_step.py --steps 4 2>&1 | tail -2
