"""bench.py contract checks that need no GPU: the reference (CPU) arm on a tiny box prints exactly one JSON line
with the keys the driver reads, and nothing else on stdout."""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*arguments):
    command = [sys.executable, os.path.join(ROOT, "bench.py"), *arguments]
    result = subprocess.run(command, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert result.returncode == 0, result.stderr[-2000:]
    return result.stdout


def test_reference_arm_prints_one_json_line():
    stdout = run_bench("--impl", "reference", "--lattice", "8", "--steps", "2", "--warmup", "1", "--cpu-seconds", "0.2")
    lines = [line for line in stdout.splitlines() if line.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "atom-steps/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["n_gpus"] == 1 and line["steps"] == 2 and line["warmup"] == 1
    assert line["config"]["atoms"] == 512 and "workload" in line["config"]
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "atom-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["dtype"] == "f64" and line["data"] == "synthetic" and line["vs_baseline"] is None


def test_native_arm_fails_loudly_without_a_device():
    import torch

    if torch.cuda.is_available():
        return  # covered by the GPU runs
    command = [sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"]
    result = subprocess.run(command, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert result.returncode != 0
    assert "no CPU fallback" in result.stderr
    assert result.stdout.strip() == ""


def test_committed_bench_line_has_the_contract_keys():
    """The JSON line `python bench.py` printed on a B200 at the end of the round (profiles/r2z_bench_default_final.json):
    every key the driver and the judge read is there, with consistent values."""
    with open(os.path.join(ROOT, "profiles", "r2z_bench_default_final.json")) as fd:
        line = json.load(fd)
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline",
                "ms_per_step_window", "value_window", "ms_per_step_amortised", "steps_amortised"):
        assert key in line, key
    assert line["unit"] == "atom-steps/s" and line["dtype"] == "f64" and line["data"] == "synthetic"
    assert line["n_gpus"] == 1 and line["warmup"] >= 3 and line["higher_is_better"] is True and line["vs_baseline"] is None
    atoms = line["config"]["atoms"]
    assert "workload" in line["config"] and atoms >= 1_000_000
    # value is atoms * steps / time of the rebuild-representative pass; the short window is reported beside it
    assert line["ms_per_step"] == line["ms_per_step_amortised"]
    assert abs(line["value"] - atoms / (line["ms_per_step"] * 1e-3)) < 1e-6 * line["value"]
    assert line["run"]["neighbor_list"]["rebuilds_in_amortised_pass"] >= 10
    e2e = line["e2e"]
    assert e2e["unit"] == line["unit"] and e2e["h2d_bytes_per_step"] == 24 * atoms and e2e["d2h_bytes_per_step"] == 24 * atoms
    assert 0 < e2e["value"] < line["value"]
    roofline = line["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in roofline, key
    assert abs(roofline["frac"] - roofline["achieved"] / roofline["peak"]) < 1e-12
    assert roofline["kernel"] == "lj2_force_kernel" and roofline["traffic"] > 0
    baseline = line["cpu_baseline"]
    assert baseline["kind"] == "port" and baseline["cores"] >= 1 and baseline["value"] > 0 and baseline["sample"]
    # the O(N^2) cost of the reference's loop is measured over a series of sizes; the figure at the bench size is a fit
    assert len(baseline["n2_series"]) >= 5 and "FIT" in baseline["n2_fit"]["label"]
    assert line["gpu_launches"] > line["steps"]
    clocks = line["clocks"]
    assert clocks["samples"] > 0
    assert not set(clocks["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert clocks["sm_mhz"] > 0.9 * clocks["sm_max_mhz"]
    # the companion runs and the criterion-style latencies ride in the same line
    assert line["lj_1m"]["config"]["atoms"] == 1048576 and line["lj_1m"]["value"] > 0
    assert line["spce"]["config"]["ewald"]["kmax"] == 25 and line["spce"]["value"] > 0
    assert line["spce"]["roofline_extra"]["pair_kernel"]["kernel"] == "cq_force_kernel"
    assert line["spce_1m"]["config"]["atoms"] >= 1_000_000 and line["spce_1m"]["value"] > 0
    assert "move_molecule_cost" in line["criterion_us_per_call"]["water_ewald"]


def test_reference_arm_under_torchrun_prints_on_rank_zero_only():
    """The driver launches the reference arm like the native one: under torchrun with N ranks, rank 0 alone runs and
    prints the line, the other ranks exit 0 without work."""
    command = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
               "--master-port", "29731", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--lattice", "8",
               "--steps", "2", "--warmup", "1", "--cpu-seconds", "0.2"]
    result = subprocess.run(command, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert result.returncode == 0, result.stderr[-2000:]
    lines = [line for line in result.stdout.splitlines() if line.strip().startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["value"] > 0


def test_clock_sampler_waits_for_its_first_sample(tmp_path, monkeypatch):
    """`clocks.samples` must not be zero for a short timed pass: the sampler returns from start() once nvidia-smi has
    written its first line (a stand-in nvidia-smi that needs 0.3 s to attach is used here)."""
    import stat
    import time

    fake = tmp_path / "nvidia-smi"
    fake.write_text("#!/bin/bash\nsleep 0.3\nwhile true; do echo '0, 1965, 1965, 700.0, 0x0, Not Active, Not Active, Not Active, Active'; sleep 0.02; done\n")
    fake.chmod(fake.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", f"{tmp_path}:{os.environ['PATH']}")
    sys.path.insert(0, ROOT)
    import bench

    sampler = bench.ClockSampler(0)
    begin = time.perf_counter()
    sampler.start()
    assert 0.25 < time.perf_counter() - begin < 3.0
    time.sleep(0.2)
    clocks = sampler.stop()
    assert clocks["samples"] >= 5 and clocks["sm_mhz"] == 1965.0 and clocks["sm_max_mhz"] == 1965.0
    assert clocks["reasons"] == ["sw_power_cap"]
