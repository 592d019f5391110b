#!/usr/bin/env bash
# What is run on a B200 box before a round is closed (gpurun --timeout 2400 -- bash tools/gpu_round.sh):
# the GPU test suite, smoke(), the default bench line, launch lists and full captures of the dominant kernels.
# Everything lands in gpurun_out/; the summaries that are kept are copied into profiles/ afterwards.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag="${1:-r2z}"
(timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4)
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1200 python bench.py > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.err; tail -c 300 gpurun_out/${tag}_bench_default.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/${tag}_launches_lj_1M.csv python tools/profile_step.py --lattice 128x128x64 --steps 4 2>&1 | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/${tag}_launches_spce_98k.csv python tools/profile_step.py --workload spce --lattice 32 --steps 3 2>&1 | tail -1
N="ncu --set full --clock-control none --import-source on"
timeout 300 $N -k regex:ewald_rho_tiled_kernel -s 1 -c 1 -o gpurun_out/${tag}_rho python tools/profile_step.py --workload spce --lattice 32 --steps 3 2>&1 | tail -1
timeout 300 $N -k regex:ewald_force_tiled_kernel -s 1 -c 1 -o gpurun_out/${tag}_kforce python tools/profile_step.py --workload spce --lattice 32 --steps 3 2>&1 | tail -1
python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
d = json.loads(open(f"gpurun_out/{tag}_bench_default.json").read().strip().splitlines()[-1])
for key in ("main", "lj_1m", "spce", "spce_1m"):
    x = d if key == "main" else d.get(key)
    if x:
        r = x.get("roofline") or {}
        print(key, "value %.4g" % x["value"], "ms/step", x["ms_per_step"], "window", x.get("ms_per_step_window"), "frac", r.get("frac"),
              "kernel ms", r.get("avg_launch_ms") or r.get("rho_plus_force_ms"), "e2e", (x.get("e2e") or {}).get("value"))
print("clocks", d.get("clocks"), "cpu", {k: v for k, v in d.get("cpu_baseline", {}).items() if k in ("value", "cores", "kind")})
PY
