"""lumol_b200: B200-native (CUDA sm_100a, FP64) force evaluation behind lumol's interfaces.

The package is the host side of the drop-in: Python mirrors of the reference's ``System`` /
``PairInteraction`` / ``Ewald`` / ``Compute`` / ``MolecularDynamics`` surface (same names, argument
meaning and error behaviour) over the C ABI of ``liblumol_cuda.so`` (``include/lumol_cuda.h``).
All per-pair, per-k-vector and per-atom arithmetic runs in hand-written CUDA kernels; without the
built library or without a CUDA device every evaluation raises ``LumolCudaError``.
"""

from . import consts, units
from ._ffi import LumolCudaError
from .cache import EnergyCache
from .energy import (
    BornMayerHuggins, Buckingham, CosineHarmonic, Ewald, Gaussian, Harmonic, LennardJones, Mie, Morse,
    NullPotential, PairInteraction, PairPotential, PairRestriction, Potential, SharedEwald, TableComputation,
    Torsion, Wolf,
)
from .sys import Molecule, Particle, System, UnitCell, system_from_xyz

__all__ = [
    "consts", "units", "LumolCudaError", "EnergyCache", "BornMayerHuggins", "Buckingham", "CosineHarmonic", "Ewald", "Gaussian",
    "Harmonic", "LennardJones", "Mie", "Morse", "NullPotential", "PairInteraction", "PairPotential",
    "PairRestriction", "Potential", "SharedEwald", "TableComputation", "Torsion", "Wolf", "Molecule", "Particle",
    "System", "UnitCell", "system_from_xyz",
]
