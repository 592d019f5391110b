cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_bench_parity.py -q -m gpu -k "lj_bench" 2>&1 | tail -4
(time timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err) 2>&1 | grep real
tail -3 gpurun_out/bench_default.err | cut -c1-300
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_default.json"))
def show(name,x):
    if "error" in x: print(name, x); return
    r=x.get("roofline") or {}
    print(name, "value %.3e ms/step %.4f window %.4f rebuilds %s frac %.3f pair_ms %s e2e %s" % (x["value"], x["ms_per_step"], x["ms_per_step_window"], x["run"]["neighbor_list"], r.get("frac") or 0, r.get("avg_launch_ms"), (x.get("e2e") or {}).get("value")))
show("main", d)
for k in ("lj_1m","spce","spce_1m"):
    if k in d: show(k, d[k])
print("clocks", d["clocks"])
print("cpu", {k:v for k,v in d["cpu_baseline"].items() if k!="sample"})
PY
