cd $GRAFT_REPO_ROOT
timeout 900 python tools/multi_device_check.py 2 2>&1 | tail -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29633 bench.py --gpus 2 --steps 400 --warmup 20 --no-spce-1m > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 400 gpurun_out/bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
for k in ('main','lj_1m','spce'):
    x=d if k=='main' else d.get(k)
    if x: print(k,'value %.4g'%x['value'],'ms',x['ms_per_step'],'e2e',x.get('e2e'))
PY
