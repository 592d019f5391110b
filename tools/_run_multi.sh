timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 300 --warmup 20 --no-e2e > gpurun_out/scale2_b.json 2> gpurun_out/scale2_b.err; echo rc=$?; python -c "
import json; d=json.loads(open('gpurun_out/scale2_b.json').read().splitlines()[-1]); print('N=2', d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], {k:v.get('ms_per_step') for k,v in d['roofline_extra'].items()}, 'spce', d['spce']['value'], d['spce']['ms_per_step'])"
wc -l gpurun_out/scale2_b.json
