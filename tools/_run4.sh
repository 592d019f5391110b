set -x
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_md.py -m gpu -x -q -k "lj or staged or large or water_box or multi_kind" 2>&1 | tail -5
timeout 200 python bench.py --steps 300 --warmup 20 --no-e2e --no-cpu-baseline > gpurun_out/bench_lj_v10.json 2> gpurun_out/bench_lj_v10.err; tail -3 gpurun_out/bench_lj_v10.err; python -c "
import json; d=json.load(open('gpurun_out/bench_lj_v10.json')); print(d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['config']['neighbor_list'], d['roofline_extra']['neighbor_list'])"
