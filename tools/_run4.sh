cd $GRAFT_REPO_ROOT
N=$1
if [ "$N" = "1" ]; then
  timeout 600 python bench.py --no-spce --no-cpu-baseline --no-e2e --steps 300 --warmup 50 > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus $N --no-spce --no-cpu-baseline --no-e2e --steps 300 --warmup 50 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
fi
tail -2 gpurun_out/scale_n$N.err | cut -c1-300
python - <<PY
import json
d=json.load(open("gpurun_out/scale_n$N.json"))
x=d["roofline_extra"]; r=d["roofline"]
print("N=$N value %.3e ms/step %.4f pair_ms %.4f rebuilds %s" % (d["value"], d["ms_per_step"], r["avg_launch_ms"], d["config"]["neighbor_list"]))
print({k:(v.get("ms_per_step")) for k,v in x.items()})
PY
