cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "lj or staged or cell_list or non_finite or edge" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_bench_parity.py -q -m gpu -k "lj_bench or cutoff" 2>&1 | tail -3
timeout 300 python bench.py --no-spce --no-cpu-baseline --no-e2e --steps 300 --warmup 50 > gpurun_out/lj2_a.json 2> gpurun_out/lj2_a.err; tail -3 gpurun_out/lj2_a.err

timeout 300 python -m pytest tests/test_gpu_md.py -q -m gpu -x 2>&1 | tail -3
python - <<'PY'
import json
for name in ("lj2_a","lj2_alllevels","lj2_nopairs"):
    try:
        d=json.load(open(f"gpurun_out/{name}.json"))
        r=d["roofline"]; x=d["roofline_extra"]
        print(name, "value %.3e ms/step %.4f pair_ms %.4f frac %.3f rebuilds %s neighbor_ms %.4f vv_ms %.4f" % (d["value"], d["ms_per_step"], r["avg_launch_ms"], r["frac"], d["config"]["neighbor_list"], x["neighbor_list"]["ms_per_step"], x["velocity_verlet"]["ms_per_step"]))
    except Exception as e:
        print(name, "failed", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 14 --csv --log-file gpurun_out/r2e_launches_lj2.csv python tools/profile_step.py --steps 2 2>&1 | tail -1

LUMOL_CUDA_SORTED_MD=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 9 --csv --log-file gpurun_out/r2e_launches_lj2b.csv python tools/profile_step.py --steps 2 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/r2e_launches_spce.csv python tools/profile_step.py --workload spce --lattice 32 --steps 2 2>&1 | tail -1
