for v in 0 1; do
if [ $v = 1 ]; then export LUMOL_CUDA_NO_REORDER=1; fi
timeout 300 python bench.py --steps 1000 --warmup 50 --no-e2e --no-cpu-baseline --no-spce > gpurun_out/bench_reorder_$v.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_reorder_$v.json')); print('noreorder=$v', d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['config']['neighbor_list'])"
done
