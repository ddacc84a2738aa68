"""bench.py contract checks that run without a GPU: the reference arm prints one well-formed JSON line (CPU oracle port,
the one place besides tests/ and smoke() that may execute oracle/), and our arm refuses to run without a B200."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None, timeout=900):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                          timeout=timeout, env=e, cwd=ROOT)


def test_reference_arm_prints_one_json_line():
    out = _run(["--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "1"], env={"LBC_CPU_THREADS": "8"})
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["config"]["workload"] == "config2"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == dict(value=d["value"], unit="images/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)


def test_reference_arm_only_rank0_works_under_torchrun_env():
    out = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip() == ""


def test_our_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("GPU present")
    out = _run(["--steps", "1", "--warmup", "1"])
    assert out.returncode != 0
    assert "needs a B200" in out.stderr
