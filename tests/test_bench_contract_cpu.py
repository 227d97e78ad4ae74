"""bench.py contract pieces that need no GPU: the reference arm prints ONE JSON line with the keys the driver
reads, and the product arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def test_reference_arm_line():
    p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3",
                        "--ref-budget", "6"], capture_output=True, text=True, timeout=900, cwd=str(ROOT))
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1                                   # exactly one JSON line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "messages/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] >= 1 and d["gpu_launches"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "messages/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["data"] == "synthetic" and d["dtype"] == "u8"
    # same workload wording and step accounting as the GPU arm (the driver compares them)
    sys.path.insert(0, str(ROOT))
    import bench
    assert d["config"]["workload"] == bench.WORKLOAD_C2 and d["steps"] == 2 and d["warmup"] == 3
    ref = d["cpu_baseline_reference"]
    from oracle import ref_loader
    if ref_loader.reference_available():                     # the reference's own class, timed on one core
        assert ref["kind"] == "reference" and ref["cores"] == 1 and ref["c1"]["messages"] == 1000
        assert ref["reduced_c2"]["value"] > 0 and ref["value"] == ref["reduced_c2"]["value"]
    else:
        assert ref is None


def test_reference_arm_other_ranks_exit_quietly():
    import os
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, timeout=120, cwd=str(ROOT), env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_product_arm_needs_a_device():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a device is present")
    p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1"], capture_output=True, text=True,
                       timeout=300, cwd=str(ROOT))
    assert p.returncode != 0 and "no CPU fallback" in (p.stderr + p.stdout)
