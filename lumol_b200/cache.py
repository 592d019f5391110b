"""Host-side mirror of ``lumol_core::sys::EnergyCache`` (sys/cache.rs): energy bookkeeping of Monte Carlo moves.

The reference keeps an N x N table of pair energies on the host and lets ``Ewald`` cache rho(k); here the cache is
the device-resident state itself: ``move_molecule_cost`` evaluates old and new energies of the moved molecule in one
batch of kernels (``lumol_cuda_move_molecule_cost``), ``update`` accepts the move on the device
(``lumol_cuda_move_molecule_accept``: positions and rho(k) are updated in place, nothing is re-uploaded).
As in the reference, cache integrity is left to the user: call ``init`` again after changing the system behind
the cache's back.
"""

import ctypes

import numpy as np

from . import _ffi
from .device import device_for


class EnergyCache:
    """``EnergyCache`` (cache.rs:22-138)."""

    def __init__(self):
        self._clear()
        self._updater = None

    def _clear(self):
        self.pairs = 0.0
        self.pairs_tail = 0.0
        self.bonds = 0.0
        self.angles = 0.0
        self.dihedrals = 0.0
        self.coulomb = 0.0
        self.global_ = 0.0

    def init(self, system):
        """cache.rs:72-99: one full energy evaluation; the positions stay resident for the cost calls."""
        self._clear()
        terms = device_for(system).compute(energy=True).energy
        self.pairs = terms.pairs
        self.pairs_tail = terms.pairs_tail
        self.bonds = terms.bonds
        self.angles = terms.angles
        self.dihedrals = terms.dihedrals
        self.coulomb = terms.coulomb_real + terms.coulomb_self + terms.coulomb_kspace
        self.global_ = 0.0

    def energy(self):
        """cache.rs:102-115"""
        energy = 0.0
        energy += self.pairs
        energy += self.pairs_tail
        energy += self.bonds
        energy += self.angles
        energy += self.dihedrals
        energy += self.coulomb
        energy += self.global_
        return energy

    def update(self, system):
        """cache.rs:119-128"""
        updater, self._updater = self._updater, None
        if updater is None:
            raise RuntimeError("called EnergyCache::update without call a `*_cost` function first")
        updater(self, system)

    def unused(self):
        """cache.rs:133-137"""
        self._updater = lambda cache, system: cache.init(system)

    # ---- costs -------------------------------------------------------------------------------------------
    def move_molecule_cost(self, system, molecule_id, new_positions):
        """cache.rs:145-213.  ``system`` still holds the old positions; they are the ones resident on the device."""
        return self.move_molecules_cost(system, [molecule_id], [new_positions])[0]

    def move_molecules_cost(self, system, molecule_ids, new_positions):
        """Costs of several independent trial moves against the same state, in one batch of launches; ``update``
        accepts the first one unless ``accept(trial)`` chose another."""
        # structure, cell and coulomb settings are re-checked (and re-uploaded if they changed); the positions
        # resident on the device -- those of ``init`` plus every accepted move -- are kept
        # ... except after move_all_molecules_cost: that call left its TRIAL configuration on the device (the reference
        # evaluates it on the same, temporarily modified System and swaps back on rejection), so the host positions, which
        # are the truth whether the trial was accepted or rejected, are uploaded again
        device = device_for(system, positions=system._device is None or getattr(self, "_trial_resident", False))
        self._trial_resident = False
        ids = np.ascontiguousarray(molecule_ids, dtype=np.int64)
        for molecule_id, positions in zip(ids, new_positions):
            bonding = system.molecule(int(molecule_id))
            if np.shape(positions) != (bonding.size(), 3):
                raise ValueError("new_positions must hold one position per particle of the molecule")
        flat = np.ascontiguousarray(np.concatenate([np.asarray(p, dtype=np.float64).reshape(-1, 3) for p in new_positions]))
        costs = (_ffi.Energy * len(ids))()
        _ffi.check(device.ctx, device.lib.lumol_cuda_move_molecules_cost(
            device.ctx, len(ids), ids.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), _ffi.as_double_pointer(flat), costs))
        pairs_delta = [c.pairs for c in costs]
        coulomb_delta = [c.coulomb_real + c.coulomb_kspace for c in costs]
        self._accepted = 0

        def updater(cache, system):
            trial = cache._accepted
            cache.pairs += pairs_delta[trial]
            cache.coulomb += coulomb_delta[trial]
            _ffi.check(device.ctx, device.lib.lumol_cuda_move_molecule_accept(device.ctx, trial))

        self._updater = updater
        return [p + c for p, c in zip(pairs_delta, coulomb_delta)]

    def accept(self, trial):
        """Which trial of the last ``move_molecules_cost`` batch ``update`` applies."""
        self._accepted = int(trial)

    def move_all_molecules_cost(self, system):
        """cache.rs:230-283: ``system`` is the system after every molecule moved rigidly (possibly in a new cell).
        Everything is re-evaluated, as in the reference ("temporarily, recompute all interactions"); the pair sum
        includes the intra-molecular pairs, which a rigid move leaves unchanged."""
        terms = device_for(system).compute(energy=True, parts=_ffi.PART_PAIRS | _ffi.PART_COULOMB).energy
        self._trial_resident = True  # the device now holds the trial positions and cell, not necessarily the accepted state
        pairs, pairs_tail = terms.pairs, terms.pairs_tail
        coulomb = terms.coulomb_real + terms.coulomb_self + terms.coulomb_kspace
        cost = (pairs - self.pairs) + (pairs_tail - self.pairs_tail) + (coulomb - self.coulomb)

        def updater(cache, system):
            cache.pairs = pairs
            cache.pairs_tail = pairs_tail
            cache.coulomb = coulomb

        self._updater = updater
        return cost
