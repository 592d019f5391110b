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
