import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import systems
from lumol_b200 import md
from oracle import oracle
s = systems.argon()
print("masses", s.masses[:3], "E", s.potential_energy())
systems.random_velocities(s, 120.0, seed=1)
print("v", np.abs(s.velocities).max())
o = oracle.OracleSystem(s)
print("K gpu", s.kinetic_energy(), "K ref", o.kinetic_energy(), "T", s.temperature())
print("E", s.potential_energy(), o.potential_energy())
p = md.MolecularDynamics(1.0)
for k in range(5):
    p.propagate(s, 1)
    print(k, "K", s.kinetic_energy(), "U", s.potential_energy(), "maxv", np.abs(s.velocities).max(), "maxx", np.abs(s.positions).max())
s = systems.propane()
o = oracle.OracleSystem(s)
f = s.forces(); fr = o.forces()
print("propane force err", np.abs(f-fr).max()/np.abs(fr).max())
from lumol_b200.device import device_for
from lumol_b200 import _ffi
d = device_for(s)
for parts in (1,2,4):
    r = d.compute(forces=True, energy=True, parts=parts)
    print(parts, r.energy.pairs, r.energy.bonds, r.energy.angles, r.energy.dihedrals, np.abs(r.forces).max())
t = o.energy_terms(); print(t.pairs, t.bonds, t.angles, t.dihedrals)
fb = np.zeros_like(fr); o.lib.orc_bonded_forces(o.ref, oracle.dptr(fb)); print("ref bonded max", np.abs(fb).max(), "pair max", np.abs(o.pair_forces()).max())
r = d.compute(forces=True, parts=2); print("bonded err", np.abs(r.forces-fb).max())
