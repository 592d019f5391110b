#!/usr/bin/env python
"""Benchmark of the force-evaluation hot path: atom-steps/s and ns/day of NVE velocity-Verlet MD.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload lj|spce] [--atoms-side S] [--impl reference]

One JSON line on stdout (rank 0).  ``value`` is the whole-job throughput of K device-resident MD steps
(positions, velocities, forces never leave HBM); ``e2e`` is the same force evaluation driven the way lumol's
host-side integrator would drive it through the C ABI: every step uploads the positions from pinned host
memory and downloads the forces.  ``roofline`` describes the dominant kernel (the pair kernel), timed live
with CUDA events on the library's stream in a separate profiled pass; ``cpu_baseline`` is the CPU oracle
(a restatement of lumol's O(N^2) rayon loop, OpenMP on every host core) on a bounded sample of the same
workload.  ``--impl reference`` prints that CPU arm alone.
"""

import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "atom-steps/sec (LJ argon NVE velocity-Verlet MD, FP64)"
UNIT = "atom-steps/s"
TIMESTEP_FS = 1.0
# algorithmic work per unit (restated in DESIGN.md)
FLOP_PER_LJ_PAIR_FORCE = 36.0
FLOP_PER_COULOMB_PAIR = 75.0
FLOP_PER_ATOM_K_RHO = 16.0
FLOP_PER_ATOM_K_FORCE = 21.0
BYTES_PER_ATOM_VV = 128.0  # merged second-half + first-half kick and drift (SURVEY 8d counts 208 B for the two separate passes)
BYTES_PER_ATOM_SORT = 100.0
# The largest single-GPU configuration of BASELINE.json configs[4] ("synthetic 1M-8M-atom LJ ... boxes"): 8 388 608 atoms.
# The 1 048 576-atom box of round 1 runs beside it ("lj_1m" in the JSON line) at every N.
DEFAULT_LATTICE = "256x256x128"


def parse_args():
    parser = argparse.ArgumentParser()
    parser.add_argument("--gpus", type=int, default=1)
    parser.add_argument("--steps", type=int, default=1000)
    parser.add_argument("--warmup", type=int, default=50)
    parser.add_argument("--impl", default="native", choices=["native", "reference"])
    parser.add_argument("--workload", default="lj", choices=["lj", "spce"])
    parser.add_argument("--lattice", default=DEFAULT_LATTICE, help="lattice points per axis (lj) or molecules per axis (spce)")
    parser.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    parser.add_argument("--no-cpu-baseline", action="store_true")
    parser.add_argument("--no-e2e", action="store_true")
    parser.add_argument("--no-spce", action="store_true", help="skip the SPC/E + Ewald companion run of the default (lj) bench")
    parser.add_argument("--spce-lattice", default="32", help="molecules per axis of the SPC/E companion run")
    parser.add_argument("--spce-steps", type=int, default=40)
    parser.add_argument("--no-lj-1m", action="store_true", help="skip the 1 048 576-atom companion run of the default (lj) bench")
    parser.add_argument("--no-spce-1m", action="store_true", help="skip the 1 029 000-atom SPC/E point (N = 1 only)")
    parser.add_argument("--amortise-steps", type=int, default=0, help="steps of the rebuild-representative pass (default: max(400, 10 K))")
    return parser.parse_args()


def lattice_of(text):
    parts = [int(p) for p in text.lower().split("x")]
    if len(parts) == 1:
        parts = parts * 3
    return tuple(parts)


def build_workload(args, workload=None, lattice_text=None):
    """The synthetic boxes of SURVEY section 8d, through the host-side API."""
    from lumol_b200 import synthetic
    import lumol_b200 as lumol

    workload = workload or args.workload
    lattice_text = lattice_text or args.lattice
    lattice = lattice_of(lattice_text)
    if workload == "lj":
        system = synthetic.lj_box(lattice, seed=20240 + 20)
        synthetic.maxwell_boltzmann(system, 120.0, seed=7)
        description = {
            "workload": f"synthetic LJ argon box, {system.size()} atoms, lattice {lattice_text}, rho=0.0213/A^3, "
                        "sigma=3.4 A, eps=1 kJ/mol, rc=10 A, tail corrections, NVE velocity-Verlet dt=1 fs, 120 K",
            "atoms": system.size(),
            "cell_A": [round(float(v), 3) for v in system.cell.lengths()],
        }
    else:
        if len(set(lattice)) != 1:
            raise SystemExit("--workload spce needs a cubic lattice")
        system = synthetic.spce_box(lattice[0], flexible=True)
        ewald = lumol.Ewald.with_accuracy(9.0, 1e-5, system)
        shared = lumol.SharedEwald(ewald)
        shared.set_restriction(lumol.PairRestriction.InterMolecular)
        system.set_coulomb_potential(shared)
        synthetic.maxwell_boltzmann(system, 300.0, seed=7)
        description = {
            "workload": f"synthetic flexible SPC/E water box, {system.size()} atoms, {lattice[0]}^3 molecules, "
                        f"O-O LJ rc=9 A + Ewald rc=9 A alpha={ewald.alpha:.4f} kmax={ewald.kmax} (with_accuracy 1e-5), "
                        "inter-molecular, harmonic bonds/angles, NVE velocity-Verlet dt=1 fs, 300 K",
            "atoms": system.size(),
            "cell_A": [round(float(v), 3) for v in system.cell.lengths()],
            "ewald": {"alpha": ewald.alpha, "kmax": ewald.kmax},
        }
    return system, description


# ---- clocks ------------------------------------------------------------------------------------------------

class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.process = None

    def start(self):
        try:
            self.process = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.device)],
                stdout=self.file, stderr=subprocess.DEVNULL,
            )
        except OSError:
            self.process = None
            return
        # nvidia-smi needs a moment to attach to the driver: the timed pass starts once the first sample is on disk
        deadline = time.perf_counter() + 5.0
        while time.perf_counter() < deadline and os.path.getsize(self.file.name) == 0 and self.process.poll() is None:
            time.sleep(0.02)

    def stop(self):
        if self.process is not None:
            self.process.terminate()
            try:
                self.process.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.process.kill()
        self.file.flush()
        self.file.seek(0)
        clocks, max_clock, reasons = [], None, set()
        for line in self.file.read().splitlines():
            fields = [f.strip() for f in line.split(",")]
            if len(fields) < 9:
                continue
            try:
                clocks.append(float(fields[1]))
                max_clock = float(fields[2])
            except ValueError:
                continue
            for name, value in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), fields[5:9]):
                if value.lower().startswith("active"):
                    reasons.add(name)
        self.file.close()
        os.unlink(self.file.name)
        if not clocks:
            return {"sm_mhz": None, "sm_max_mhz": max_clock, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(clocks)), "sm_max_mhz": max_clock, "reasons": sorted(reasons), "samples": len(clocks)}


# ---- CPU arm -------------------------------------------------------------------------------------------------

class CpuSampler:
    """The oracle (restatement of lumol's pair-force loop, sys/compute.rs:37-55, OpenMP on every host core) timed on
    uniformly spaced rows i of the O(N^2) loop.  Test infrastructure used as the baseline."""

    def __init__(self, system):
        from oracle import oracle

        self.oracle = oracle
        self.reference = oracle.OracleSystem(system)
        self.threads = os.cpu_count() or 1
        self.reference.lib.orc_set_threads(self.threads)
        self.n = system.size()
        self.checksum = ctypes.c_double()

    def run(self, count):
        rows = np.ascontiguousarray(np.linspace(0, self.n - 1, count).astype(np.int64))
        start = time.perf_counter()
        self.reference.lib.orc_pair_forces_sample(self.reference.ref, len(rows), self.oracle.iptr(rows), ctypes.byref(self.checksum))
        return time.perf_counter() - start

    def calibrate(self, target_seconds):
        """Number of rows that runs for about ``target_seconds`` (the first probes are dominated by thread start-up)."""
        rows = min(self.n, max(self.threads * 2, 16))
        seconds = self.run(rows)
        while seconds < 0.6 * target_seconds and rows < self.n:
            rows = int(min(self.n, max(rows * 2, rows * target_seconds / max(seconds, 1e-3))))
            rows = max(self.threads, rows // self.threads * self.threads)
            seconds = self.run(rows)
        return rows, seconds


def cpu_sample_rows(system, target_seconds):
    """One bounded sample: (rows per second, rows, seconds, threads)."""
    sampler = CpuSampler(system)
    rows, seconds = sampler.calibrate(target_seconds)
    return rows / seconds, rows, seconds, sampler.threads


N2_SERIES_SIDES = (7, 10, 16, 25, 40)  # 343 ... 64 000 atoms: full O(N^2) evaluations of the same LJ fluid


def cpu_n2_series(threads):
    """SURVEY 8d / BASELINE.md: the reference's pair-force loop (compute.rs:37-55, restated in the oracle) timed in FULL on
    boxes the CPU can finish, and the N^2 law fitted to them; the figure quoted for the bench box is this fit, labelled so."""
    from lumol_b200 import synthetic
    from oracle import oracle

    series = []
    for side in N2_SERIES_SIDES:
        system = synthetic.lj_box(side, seed=20240 + side)
        reference = oracle.OracleSystem(system)
        reference.lib.orc_set_threads(threads)
        reference.pair_forces()  # threads and buffers warm
        start = time.perf_counter()
        reference.pair_forces()
        series.append({"atoms": system.size(), "seconds": time.perf_counter() - start})
    # least squares of t = a N^2 through the three largest sizes (thread start-up dominates the small ones)
    tail = series[-3:]
    a = sum(p["seconds"] * p["atoms"] ** 2 for p in tail) / sum(float(p["atoms"]) ** 4 for p in tail)
    return series, a


def cpu_baseline_for(system, target_seconds):
    """cpu_baseline of the JSON line: measured rows of the bench box itself, and the measured N^2 series with its fit."""
    n = system.size()
    rate, rows, seconds, threads = cpu_sample_rows(system, target_seconds)
    series, a = cpu_n2_series(threads)
    fit_seconds = a * float(n) ** 2
    return {
        "value": rate, "unit": UNIT, "cores": threads, "kind": "port",
        "sample": f"pair-force rows of {rows} of {n} atoms (uniform stride), all j > i: {seconds:.1f} s of the O(N^2) "
                  "loop of sys/compute.rs:37-55 restated in oracle/lumol_oracle.c (OpenMP); value = rows per second, i.e. the "
                  "rate of a full evaluation as sampled on this very box (forces only, no integration)",
        "n2_series": series,
        "n2_fit": {"seconds_per_evaluation": f"{a:.4e} * N^2", "seconds_at_bench_size": fit_seconds,
                   "atom_steps_per_s_at_bench_size": n / fit_seconds,
                   "label": "FIT of the measured series above, not a measurement at the bench size"},
    }


def run_reference(args):
    """``--impl reference``: the CPU path alone, same metric and config: W + K bounded samples ("steps") of the same
    size, sized once so that the whole run takes about a minute and a half whatever K and W are."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    system, description = build_workload(args)
    if args.workload != "lj":
        description["note"] = "CPU arm times the pair-force loop only"
    count = max(args.steps + args.warmup, 1)
    per_step = min(args.cpu_seconds, 90.0 / count)
    sampler = CpuSampler(system)
    rows, _ = sampler.calibrate(per_step)
    values = []
    for step in range(args.warmup + args.steps):
        seconds = sampler.run(rows)
        if step >= args.warmup:
            values.append(rows / seconds)
    value = float(np.mean(values)) if values else 0.0
    n = system.size()
    threads = sampler.threads
    sample = (f"per step: pair-force rows of {rows} of {n} atoms (uniform stride), all j > i, O(N^2) loop of "
              "sys/compute.rs:37-55 restated in oracle/lumol_oracle.c (OpenMP); value = rows per second = the rate of full "
              "evaluations as sampled on this box (forces only, no integration)")
    baseline = {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}
    if args.workload == "lj" and n >= 100000:
        series, a = cpu_n2_series(threads)
        fit_seconds = a * float(n) ** 2
        baseline["n2_series"] = series
        baseline["n2_fit"] = {"seconds_per_evaluation": f"{a:.4e} * N^2", "seconds_at_bench_size": fit_seconds,
                              "atom_steps_per_s_at_bench_size": n / fit_seconds,
                              "label": "FIT of the measured series above, not a measurement at the bench size"}
    emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * n / value if value else None, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": description,
        "cpu_baseline": baseline,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "ns_per_day": value / n * TIMESTEP_FS * 86400.0 * 1e-6 if value else None,
    }))


# ---- GPU arm ---------------------------------------------------------------------------------------------------

def bench_system_latencies(repeats=30):
    """Per-call latency (microseconds, host API, host buffers in and out) of forces / potential energy / atomic
    virial on the reference's four criterion bench systems (benches/{argon,nacl,water,propane}.rs; 128-300 atoms,
    SURVEY section 6), for continuity with the reference's own harness."""
    tests = os.path.join(ROOT, "tests")
    if tests not in sys.path:
        sys.path.insert(0, tests)
    import systems
    from lumol_b200.compute import AtomicVirial, Forces, PotentialEnergy

    out = {}
    for name, builder in (("argon", systems.argon), ("nacl_ewald", lambda: systems.nacl("ewald")),
                          ("nacl_wolf", lambda: systems.nacl("wolf")), ("water_ewald", lambda: systems.water("ewald")),
                          ("propane", systems.propane)):
        system = builder()
        row = {"atoms": system.size()}
        for label, estimator in (("forces", Forces()), ("energy", PotentialEnergy()), ("virial", AtomicVirial())):
            for _ in range(3):
                estimator.compute(system)
            start = time.perf_counter()
            for _ in range(repeats):
                estimator.compute(system)
            row[label] = (time.perf_counter() - start) / repeats * 1e6
        try:
            # benches/*.rs `move_molecule_cost` / `move_all_molecules_cost`: one trial per call, and 64 trials per batch
            import lumol_b200 as lumol

            rng = np.random.Generator(np.random.PCG64(17))
            cache = lumol.EnergyCache()
            cache.init(system)
            nmol = len(system.molecules())

            def trial():
                molecule = int(rng.integers(0, nmol))
                bonding = system.molecule(molecule)
                return molecule, system.positions[bonding.start:bonding.end] + rng.uniform(-1.0, 1.0, 3)

            trials = [trial() for _ in range(64)]
            for molecule, positions in trials[:3]:
                cache.move_molecule_cost(system, molecule, positions)
            start = time.perf_counter()
            for molecule, positions in trials[:repeats]:
                cache.move_molecule_cost(system, molecule, positions)
            row["move_molecule_cost"] = (time.perf_counter() - start) / repeats * 1e6
            ids, news = [t[0] for t in trials], [t[1] for t in trials]
            cache.move_molecules_cost(system, ids, news)
            start = time.perf_counter()
            for _ in range(5):
                cache.move_molecules_cost(system, ids, news)
            row["move_molecule_cost_batch64_per_trial"] = (time.perf_counter() - start) / (5 * 64) * 1e6
            start = time.perf_counter()
            for _ in range(5):
                cache.move_all_molecules_cost(system)
            row["move_all_molecules_cost"] = (time.perf_counter() - start) / 5 * 1e6
        except Exception as error:  # the headline numbers must survive a failure of this extra
            row["move_molecule_cost_error"] = str(error)
        system._device.close()
        out[name] = row
    # device-resident MD of the smallest system (300 atoms): launch-latency bound, replayed from a CUDA graph
    from lumol_b200 import md

    system = systems.lj_box(7, seed=3)  # 343 atoms, all-pairs path (argon.pdb holds overlapping atoms: not for MD)
    systems.random_velocities(system, 120.0, seed=1)
    propagator = md.MolecularDynamics(TIMESTEP_FS)
    propagator.propagate(system, 100, download=False)
    start = time.perf_counter()
    propagator.propagate(system, 5000, download=False)
    out["md_343_atoms_us_per_step"] = (time.perf_counter() - start) / 5000 * 1e6
    system._device.close()
    return out


def load_json(name):
    try:
        with open(os.path.join(ROOT, name)) as fd:
            return json.load(fd)
    except (OSError, ValueError):
        return {}


def measure(args, env, workload, lattice_text, steps, warmup, with_e2e, with_cpu, amortise_steps=None):
    """One workload on this rank's GPU: device-resident MD (value), e2e through the C ABI with host buffers, and a
    profiled pass for the rooflines.  Returns the result dict on rank 0, None elsewhere.

    Three timed passes of device-resident MD, all with CUDA events on the library's stream, max over ranks:
      * the window: exactly ``steps`` steps after ``warmup`` (what the driver asks for);
      * the rebuild-representative pass: ``amortise_steps`` steps (default max(400, 10 K)), long enough to hold several
        neighbour-list rebuilds; ``value`` is the throughput of THIS pass (a 20-step window usually holds no rebuild and
        would overstate the sustained rate);
      * the profiled pass (events around every kernel class, host synchronisation in between): per-launch times for the
        rooflines only.
    """
    import torch
    import torch.distributed as dist

    from lumol_b200 import _ffi, md, parallel
    from lumol_b200.device import device_for

    rank, world, local_rank = env
    system, description = build_workload(args, workload, lattice_text)
    system.device_ordinal = local_rank
    n = system.size()
    device = device_for(system, velocities=True)
    lib, ctx = device.lib, device.ctx
    if world > 1:
        parallel.init_communicator(device, rank, world)
    stream = torch.cuda.ExternalStream(lib.lumol_cuda_stream(ctx))

    def barrier():
        lib.lumol_cuda_synchronize(ctx)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(milliseconds):
        return parallel.max_over_ranks(milliseconds, world)

    def timed_run(count):
        """(milliseconds, rebuilds, kernel launches) of ``count`` device-resident MD steps."""
        before = device.stats()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        start.record(stream)
        _ffi.check(ctx, lib.lumol_cuda_md_run(ctx, count))
        stop.record(stream)
        barrier()
        after = device.stats()
        return (max_over_ranks(start.elapsed_time(stop)), int(after.neighbor_rebuilds - before.neighbor_rebuilds),
                int(after.kernel_launches - before.kernel_launches))

    propagator = md.MolecularDynamics(TIMESTEP_FS)
    propagator.setup(system)

    # ---- device-resident MD: warm-up, then exactly K timed steps, then the rebuild-representative pass -------------
    _ffi.check(ctx, lib.lumol_cuda_md_run(ctx, warmup))
    barrier()
    _ffi.check(ctx, lib.lumol_cuda_reset_stats(ctx))
    window_ms, rebuilds_window, launches = timed_run(steps)
    long_steps = amortise_steps if amortise_steps else (args.amortise_steps or max(400, 10 * steps))
    if not args.amortise_steps:
        # at least ~2.5 s, so that the 20 ms clock samples of this pass are a population and not a handful
        long_steps = min(max(long_steps, int(2500.0 / max(window_ms / max(steps, 1), 1e-3))), 20000)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    long_ms, rebuilds_long, _ = timed_run(long_steps)
    clocks = sampler.stop() if rank == 0 else None
    stats = device.stats()
    value = n * long_steps / (long_ms * 1e-3)

    # ---- end to end through the C ABI with host buffers -------------------------------------------------------
    e2e = None
    if with_e2e:
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        host_positions = torch.empty((n, 3), dtype=torch.float64, pin_memory=True)
        host_forces = torch.empty((n, 3), dtype=torch.float64, pin_memory=True)
        positions = np.zeros((n, 3))
        _ffi.check(ctx, lib.lumol_cuda_get_positions(ctx, _ffi.as_double_pointer(positions)))
        host_positions.copy_(torch.from_numpy(positions))
        p_in = ctypes.cast(host_positions.data_ptr(), ctypes.POINTER(ctypes.c_double))
        p_out = ctypes.cast(host_forces.data_ptr(), ctypes.POINTER(ctypes.c_double))
        e2e_steps = max(3, min(steps, 20 if n > 2_000_000 else 50))

        # sharded: every rank uploads the positions of the block of atoms its host-side integrator advances (the ranks
        # exchange the blocks on the device over NVLink) and reads back the forces of that block; the union of the blocks
        # is the state of the run
        first, owned = ctypes.c_int64(0), ctypes.c_int64(n)
        if world > 1:
            _ffi.check(ctx, lib.lumol_cuda_owned_range(ctx, ctypes.byref(first), ctypes.byref(owned)))
        what = _ffi.FORCES | (_ffi.OWNED_FORCES if world > 1 else 0)
        p_owned = ctypes.cast(host_positions.data_ptr() + 24 * int(first.value), ctypes.POINTER(ctypes.c_double))

        def e2e_step():
            # what lumol's VelocityVerlet::integrate does around system.forces() (integrators.rs:44-69) when the
            # host arrays are the truth: positions in, forces out
            if world > 1:
                _ffi.check(ctx, lib.lumol_cuda_set_owned_positions(ctx, p_owned))
            else:
                _ffi.check(ctx, lib.lumol_cuda_set_positions(ctx, p_in))
            _ffi.check(ctx, lib.lumol_cuda_compute(ctx, what, _ffi.PART_ALL, p_out, None, None))

        for _ in range(3):
            e2e_step()
        barrier()
        start.record(stream)
        for _ in range(e2e_steps):
            e2e_step()
        stop.record(stream)
        barrier()
        e2e_ms = max_over_ranks(start.elapsed_time(stop))
        e2e = {
            # bytes of the whole job (all ranks together); each rank moves its own block over its own PCIe link
            "value": n * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 24 * n, "d2h_bytes_per_step": 24 * n,
            "bytes_per_step_per_rank": 24 * int(owned.value),
            "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
            "call": ("lumol_cuda_set_positions + lumol_cuda_compute(FORCES) with pinned host buffers" if world == 1 else
                     "per rank: lumol_cuda_set_owned_positions (own block up, blocks exchanged over NVLink) + "
                     "lumol_cuda_compute(FORCES | OWNED_FORCES) (own block down), pinned host buffers"),
        }
        # restore the device-resident state of the MD run
        _ffi.check(ctx, lib.lumol_cuda_set_positions(ctx, _ffi.as_double_pointer(positions)))

    # ---- roofline of the dominant kernels: separate profiled pass (events around each kernel class) -------------
    energy = _ffi.Energy()
    _ffi.check(ctx, lib.lumol_cuda_compute(ctx, _ffi.ENERGY, _ffi.PART_ALL, None, ctypes.byref(energy), None))
    counts = device.stats()
    pair_count, coulomb_pairs, nk = counts.pair_count, counts.coulomb_pair_count, int(counts.nkvectors)
    fp64_peak = ctypes.c_double()
    _ffi.check(ctx, lib.lumol_cuda_measure_fp64_peak(ctx, ctypes.byref(fp64_peak)))
    profile_steps = max(3, min(steps, 20))
    propagator.setup(system)
    _ffi.check(ctx, lib.lumol_cuda_md_run(ctx, 2))
    _ffi.check(ctx, lib.lumol_cuda_reset_stats(ctx))
    rebuilds_before_profile = int(device.stats().neighbor_rebuilds)
    _ffi.check(ctx, lib.lumol_cuda_set_profiling(ctx, 1))
    _ffi.check(ctx, lib.lumol_cuda_md_run(ctx, profile_steps))
    _ffi.check(ctx, lib.lumol_cuda_set_profiling(ctx, 0))
    profile = device.stats()
    barrier()
    device.close()

    if rank != 0:
        return None

    peaks = load_json("MEASURED_PEAKS.json")
    traffic = load_json(os.path.join("profiles", "traffic.json"))
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_source = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
    fp64_source = ("measured live: dependent-free DFMA chains on every SM (lumol_cuda_measure_fp64_peak); "
                   "MEASURED_PEAKS.json has no FP64 entry")

    def per_launch(clock_ms, clock_launches):
        return clock_ms / clock_launches if clock_launches else None

    on_list = counts.neighbor_path == 1
    if on_list and workload == "lj":
        pair_kernel = "lj2_force_kernel"
    elif on_list:
        pair_kernel = "cq_force_kernel"
    else:
        pair_kernel = "allpairs_kernel"
    pair_ms = per_launch(profile.pair_ms, profile.pair_launches)
    pair_flops = (pair_count * FLOP_PER_LJ_PAIR_FORCE + coulomb_pairs * FLOP_PER_COULOMB_PAIR) / world
    kspace_ms = per_launch(profile.kspace_ms, profile.kspace_launches)
    roofline_pair = None
    if pair_ms:
        achieved = pair_flops / (pair_ms * 1e-3) / 1e12
        entry = traffic.get(f"{pair_kernel}:{workload}:{n}") or {}
        roofline_pair = {
            "kernel": pair_kernel, "bound": "fp64", "achieved": achieved, "peak": fp64_peak.value, "unit": "TFLOP/s",
            "frac": achieved / fp64_peak.value, "peak_source": fp64_source,
            "algorithmic_flop_per_launch": pair_flops, "pairs_in_cutoff": pair_count, "coulomb_pairs_in_cutoff": coulomb_pairs,
            "avg_launch_ms": pair_ms, "launches_timed": int(profile.pair_launches),
            "traffic": entry.get("bytes"), "traffic_source": entry.get("source"),
        }
    roofline_extra = {}
    if kspace_ms and nk:
        # one rho launch + one force launch per evaluation
        flops = n * nk * (FLOP_PER_ATOM_K_RHO + FLOP_PER_ATOM_K_FORCE) / world
        total_ms = profile.kspace_ms / (profile.kspace_launches / 2.0)
        achieved = flops / (total_ms * 1e-3) / 1e12
        issued = achieved * 12.0 / (FLOP_PER_ATOM_K_RHO + FLOP_PER_ATOM_K_FORCE)
        roofline_extra["ewald_kspace"] = {
            "kernel": "ewald_rho_tiled_kernel + ewald_force_tiled_kernel", "bound": "fp64",
            # the tiled kernels issue 12 FLOP per (atom, k) pair (2 + 4 DFMA: the +l / -l products are shared) where the
            # reference's arithmetic, which SURVEY 8d counts, needs 37: `frac` is the issued rate against the DFMA peak
            "achieved": issued, "peak": fp64_peak.value, "unit": "TFLOP/s", "frac": issued / fp64_peak.value,
            "peak_source": fp64_source, "nkvectors": nk, "rho_plus_force_ms": total_ms,
            "reference_flop_per_evaluation": flops, "reference_arithmetic_tflops": achieved,
            "note": "achieved / frac count the FLOP the kernels issue (12 per atom-k pair); by the reference's own arithmetic "
                    "(16 + 21 FLOP per pair, SURVEY 8d) the same time corresponds to reference_arithmetic_tflops",
            "traffic": (traffic.get(f"ewald_kspace:{workload}:{n}") or {}).get("bytes"),
            "traffic_source": (traffic.get(f"ewald_kspace:{workload}:{n}") or {}).get("source"),
        }
    if pair_ms and roofline_pair is not None and "ewald_kspace" in roofline_extra:
        roofline_extra["pair_kernel"] = roofline_pair
    if profile.integrate_launches:
        total_ms = profile.integrate_ms / profile_steps
        # sorted-resident engine: kick + drift + frame refresh in one pass (x, v, f, m, reference position in; x, v and
        # the frame images out): 24 * 4 + 8 + 24 + 8 read, 24 * 2 + 29 written
        per_atom = 213.0 if pair_kernel == "lj2_force_kernel" else BYTES_PER_ATOM_VV
        achieved = per_atom * n / world / (total_ms * 1e-3) / 1e9
        roofline_extra["velocity_verlet"] = {
            "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
            "peak_source": hbm_source, "ms_per_step": total_ms, "bytes_per_atom_step": per_atom,
            "what": "kick + drift + refresh of the cell-ordered frames and displacement check, one kernel" if per_atom > 200
                    else "merged second-half + first-half kick and drift",
        }
    if profile.neighbor_launches:
        total_ms = profile.neighbor_ms / profile_steps
        roofline_extra["neighbor_list"] = {
            "ms_per_step": total_ms,
            "rebuilds_in_profiled_steps": int(profile.neighbor_rebuilds - rebuilds_before_profile),
            "skin_A": profile.neighbor_skin,
            "what": "rebuild guards (two launches that return at once) plus the rebuilds that fell into the profiled steps",
        }
    if profile.comm_launches:
        roofline_extra["collectives"] = {
            "ms_per_step": profile.comm_ms / profile_steps, "launches_per_step": profile.comm_launches / profile_steps,
            "what": ("halo frames pushed to the peers that stage them (NVLink peer stores), arrival stamps and rebuild decision, "
                     "device-side all-gather at rebuilds; the profiled pass adds host synchronisation between the kernel classes, so "
                     "the waits for the other ranks are longer here than in the timed passes") if pair_kernel == "lj2_force_kernel"
                    else "ncclAllGather of the positions (24 B/atom) after the drift; ncclAllReduce of rho(k) with Ewald",
        }
    dominant = roofline_pair
    if "ewald_kspace" in roofline_extra and kspace_ms and pair_ms and profile.kspace_ms > profile.pair_ms:
        dominant = roofline_extra["ewald_kspace"]

    cpu_baseline = None
    if world == 1 and with_cpu:
        cpu_baseline = cpu_baseline_for(system, args.cpu_seconds)

    run = {
        "parallelism": (f"{world} x B200: units of 256 cell-ordered atoms owned by ranks, halo frames over NVLink peer memory"
                        if pair_kernel == "lj2_force_kernel" else f"{world} x B200, atoms in contiguous blocks per rank, replicated positions")
                       if world > 1 else "1 x B200",
        "neighbor_path": "cell list" if counts.neighbor_path == 1 else "all-pairs",
        "cells": [int(c) for c in counts.ncells],
        "neighbor_list": {"skin_A": counts.neighbor_skin, "rebuilds_in_window": rebuilds_window, "rebuilds_in_amortised_pass": rebuilds_long},
        "l2": f"working set (particle state, frames and a neighbour list of about {(pair_count * 2 * 1.33 * 2) / 1e6:.0f} MB) against the "
              "126 MB L2; no explicit flush",
    }
    return {
        "metric": METRIC if workload == "lj" else METRIC.replace("LJ argon NVE", "SPC/E water Ewald NVE"),
        "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        # `ms_per_step` and `value` are the rebuild-representative ones; the K-step window the driver asked for is beside them
        "ms_per_step": long_ms / long_steps, "ms_per_step_amortised": long_ms / long_steps, "steps_amortised": long_steps,
        "ms_per_step_window": window_ms / steps, "value_window": n * steps / (window_ms * 1e-3),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": description, "run": run,
        "ns_per_day": long_steps / (long_ms * 1e-3) * TIMESTEP_FS * 86400.0 * 1e-6,
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": dominant, "roofline_extra": roofline_extra,
        "cpu_baseline": cpu_baseline,
    }


_REAL_STDOUT = None


def emit(line):
    """The one JSON line of the contract, on the process's original stdout."""
    data = (line + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line + "\n")
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    args = parse_args()
    # libraries (NCCL prints its version banner) write to file descriptor 1: keep it for the JSON line alone
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: lumol_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    env = (rank, world, local_rank)

    result = measure(args, env, args.workload, args.lattice, args.steps, args.warmup, not args.no_e2e,
                     not args.no_cpu_baseline)
    keep = ("metric", "value", "unit", "steps", "warmup", "ms_per_step", "ms_per_step_window", "value_window", "steps_amortised",
            "ns_per_day", "config", "run", "e2e", "gpu_launches", "roofline", "roofline_extra")
    companions = {}
    default_bench = args.workload == "lj" and args.lattice == DEFAULT_LATTICE

    def companion(name, workload, lattice, steps, warmup, with_e2e, amortise=None):
        """A further workload in the same line; never lets the headline line be lost."""
        try:
            line = measure(args, env, workload, lattice, steps, warmup, with_e2e, False, amortise)
            if line is not None:
                companions[name] = {key: line[key] for key in keep}
        except Exception as error:
            if world > 1:
                raise  # the other ranks are inside collectives: fail together
            companions[name] = {"error": str(error)}

    if default_bench and not args.no_lj_1m:
        # the 1 048 576-atom box round 1 quoted (strong scaling of a box eight times smaller than the headline one)
        companion("lj_1m", "lj", "128x128x64", args.steps, args.warmup, not args.no_e2e)
    if args.workload == "lj" and not args.no_spce:
        # the other half of the headline metric: SPC/E water with Ewald (alpha, kmax from Ewald::with_accuracy)
        companion("spce", "spce", args.spce_lattice, args.spce_steps, max(3, min(args.warmup, 5)), not args.no_e2e, 4 * args.spce_steps)
    if default_bench and world == 1 and not args.no_spce and not args.no_spce_1m:
        # >= 1M-atom SPC/E: 70^3 molecules = 1 029 000 atoms, with_accuracy(9 A, 1e-5) -> kmax 54, 3.3e5 k-vectors, direct sum
        # like the reference's Ewald (3.4e11 atom-k products per evaluation): a few steps only
        companion("spce_1m", "spce", "70", 3, 3, False, 3)
    if rank == 0:
        if world == 1 and default_bench and not args.no_spce:
            try:
                result["criterion_us_per_call"] = bench_system_latencies()
            except Exception as error:  # an extra: never lose the JSON line to it
                result["criterion_us_per_call"] = {"error": str(error)}
        result.update(companions)
        emit(json.dumps(result))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
