timeout 900 python -m pytest tests/test_gpu_md.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
for g in 0 1; do
if [ $g = 1 ]; then export LUMOL_CUDA_NO_GRAPH=1; fi
python - <<'PY'
import sys, os, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import systems
from lumol_b200 import md
for builder, name in ((lambda: systems.lj_box(7, seed=3), 'lj343'), (systems.md_water, 'md_water'), (lambda: systems.md_nacl('ewald'), 'md_nacl_ewald')):
    s = builder(); systems.random_velocities(s, 120.0, seed=1)
    p = md.MolecularDynamics(1.0); p.propagate(s, 100, download=False)
    t = time.perf_counter(); p.propagate(s, 3000, download=False); dt = time.perf_counter() - t
    print('graph' if not os.environ.get('LUMOL_CUDA_NO_GRAPH') else 'eager', name, s.size(), 'atoms', round(dt / 3000 * 1e6, 2), 'us/step')
PY
done
python __graft_entry__.py smoke 2>&1 | tail -2
