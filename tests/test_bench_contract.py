"""CPU: the parts of bench.py's contract that do not need a GPU -- the reference arm prints one well-formed JSON
line (here with the UNMODIFIED reference when /root/reference is mounted: kind == "reference"), ranks other than 0
stay silent, and the GPU arm refuses to run without a device instead of falling back."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=timeout, env=e,
                          cwd=str(ROOT))


def test_reference_arm_prints_one_json_line():
    r = _run(["--impl", "reference", "--steps", "3", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GB/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["gpu_launches"] == 0 and d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and "median" in cb["sample"]
    from oracle import reference_loader
    assert cb["kind"] == ("reference" if reference_loader.available() else "port")
    assert d["steps"] >= 3


def test_reference_arm_other_ranks_are_silent():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "1"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("GPU present")
    r = _run(["--steps", "1", "--warmup", "1", "--no-e2e", "--no-configs", "--no-cpu-baseline"], timeout=300)
    assert r.returncode != 0 and "CUDA" in (r.stderr + r.stdout)
