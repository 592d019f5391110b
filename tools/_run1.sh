set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "lj or staged or cell_list or non_finite or edge" 2>&1 | tail -25
timeout 600 python -m pytest tests/test_gpu_bench_parity.py -q -m gpu -s -k "lj_bench or cutoff" 2>&1 | tail -25
timeout 300 python bench.py --no-spce --no-cpu-baseline --no-e2e --steps 300 --warmup 50 > gpurun_out/lj2_a.json 2> gpurun_out/lj2_a.err; tail -3 gpurun_out/lj2_a.err
LUMOL_CUDA_LJ2_ALL_LEVELS=1 timeout 300 python bench.py --no-spce --no-cpu-baseline --no-e2e --steps 300 --warmup 50 > gpurun_out/lj2_alllevels.json 2> gpurun_out/lj2_b.err
LUMOL_CUDA_LJ2=0 timeout 300 python bench.py --no-spce --no-cpu-baseline --no-e2e --steps 300 --warmup 50 > gpurun_out/lj2_off.json 2> gpurun_out/lj2_c.err
python - <<'PY'
import json
for name in ("lj2_a","lj2_alllevels","lj2_off"):
    try:
        d=json.load(open(f"gpurun_out/{name}.json"))
        r=d["roofline"]; x=d["roofline_extra"]
        print(name, "value %.3e ms/step %.4f pair_ms %.4f frac %.3f rebuilds %s neighbor_ms %.4f vv_ms %.4f" % (d["value"], d["ms_per_step"], r["avg_launch_ms"], r["frac"], d["config"]["neighbor_list"], x["neighbor_list"]["ms_per_step"], x["velocity_verlet"]["ms_per_step"]))
    except Exception as e:
        print(name, "failed", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lj2_force_kernel -s 12 -c 1 -o gpurun_out/r2d_lj2 python tools/profile_step.py --steps 16 2>&1 | tail -2
