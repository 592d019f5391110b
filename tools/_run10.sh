cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_parity.py tests/test_gpu_md.py tests/test_gpu_mc.py -q -m gpu -x 2>&1 | tail -5)
timeout 600 python bench.py --workload spce --lattice 32 --steps 40 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/spce_k2.json 2> gpurun_out/spce_k2.err; tail -c 600 gpurun_out/spce_k2.err
timeout 600 python bench.py --lattice 128x128x64 --steps 1000 --warmup 50 --no-cpu-baseline --no-e2e --no-spce --no-lj-1m --no-spce-1m > gpurun_out/lj1m_k2.json 2> gpurun_out/lj1m_k2.err; tail -c 600 gpurun_out/lj1m_k2.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/r2j_launches_lj_1M.csv python tools/profile_step.py --steps 2 2>&1 | tail -1
python - <<'PY'
import json
for f in ('gpurun_out/spce_k2.json','gpurun_out/lj1m_k2.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d.get('ms_per_step_window'), json.dumps(d.get('roofline'))[:400], json.dumps(d.get('roofline_extra'))[:600])
    except Exception as e: print(f, 'ERR', e)
PY
grep reorder gpurun_out/r2j_launches_lj_1M.csv | head -2 | cut -c1-60,200-
