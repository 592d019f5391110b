cd $GRAFT_REPO_ROOT
LUMOL_CUDA_LJ2_ALL_LEVELS=2 timeout 300 python bench.py --no-spce --no-cpu-baseline --no-e2e --steps 100 --warmup 20 > gpurun_out/lj2_nopairs.json 2> gpurun_out/lj2_d.err
python - <<'PY'
import json
for name in ("lj2_nopairs",):
    try:
        d=json.load(open(f"gpurun_out/{name}.json"))
        r=d["roofline"]; x=d["roofline_extra"]
        print(name, "value %.3e ms/step %.4f pair_ms %.4f" % (d["value"], d["ms_per_step"], r["avg_launch_ms"]))
    except Exception as e:
        print(name, "failed", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2e_launches_lj2.csv python tools/profile_step.py --steps 6 2>&1 | tail -1
LUMOL_CUDA_LJ2=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2e_launches_lj1.csv python tools/profile_step.py --steps 6 2>&1 | tail -1
