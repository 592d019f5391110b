cd $GRAFT_REPO_ROOT
N=$1
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -2 gpurun_out/bench_n$N.err | cut -c1-300
python - <<PY
import json
d=json.load(open("gpurun_out/bench_n$N.json"))
def show(name,x):
    if "error" in x: print(name, x); return
    r=x.get("roofline") or {}
    print(name, "N=$N value %.3e ms/step %.4f window %.4f rebuilds %s e2e %s" % (x["value"], x["ms_per_step"], x["ms_per_step_window"], x["run"]["neighbor_list"], (x.get("e2e") or {}).get("value")))
show("main", d)
for k in ("lj_1m","spce","spce_1m"):
    if k in d: show(k, d[k])
PY
