cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_parity.py tests/test_gpu_md.py tests/test_gpu_mc.py -q -m gpu 2>&1 | tail -15
timeout 600 python bench.py --workload spce --lattice 32 --no-cpu-baseline --no-e2e --steps 40 --warmup 5 > gpurun_out/spce_cq.json 2> gpurun_out/spce_cq.err; tail -3 gpurun_out/spce_cq.err

python - <<'PY'
import json
for name in ("spce_cq","spce_v1"):
    try:
        d=json.load(open(f"gpurun_out/{name}.json"))
        x=d["roofline_extra"]; pk=x.get("pair_kernel") or d["roofline"]
        print(name, "value %.3e ms/step %.4f pair_ms %.4f frac %.3f" % (d["value"], d["ms_per_step"], pk["avg_launch_ms"], pk["frac"]), {k:round(v.get("ms_per_step") or v.get("rho_plus_force_ms") or 0,4) for k,v in x.items()}, d["config"]["neighbor_list"])
    except Exception as e:
        print(name, "failed", e)
PY
