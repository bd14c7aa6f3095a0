"""bench.py contract (CPU tier): the reference arm prints exactly one JSON line with the keys the driver
reads; the GPU arm refuses to run without a device instead of falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=e, cwd=ROOT)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = run("--impl", "reference", "--config", "c3small", "--steps", "1", "--warmup", "1", "--ref-seconds-per-step", "0.5")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "photon primaries/sec" and d["unit"] == "primaries/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    r = run("--impl", "reference", "--config", "c3small", "--steps", "1", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_needs_a_device():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    r = run("--steps", "1", "--warmup", "1", "--config", "c3small", "--histories", "1000")
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
