"""The driver-facing contract of bench.py, as far as it can be checked without a GPU: the reference arm prints ONE
JSON line with the agreed keys, and our arm refuses to run (no CPU fallback) when no device is present."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout=900):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                          timeout=timeout, cwd=ROOT)


@pytest.mark.timeout(1200)
def test_reference_arm_prints_one_contract_line():
    r = _run(["--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "tok/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("Mistral-7B decode tok/s") and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["value"] > 0 and abs(d["ms_per_step"] * d["value"] - 1000.0) < 1e-3 * 1000.0
    assert d["config"]["workload"].startswith("Mistral-7B FP16 single-stream decode") and d["config"]["effort"] == 0.25
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None and d["data"] == "synthetic"


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_our_arm_needs_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    r = _run(["--gpus", "1", "--steps", "1", "--warmup", "1", "--quick", "--no-cpu"], timeout=300)
    assert r.returncode != 0                      # fails loudly: there is no CPU path behind the product API
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]
