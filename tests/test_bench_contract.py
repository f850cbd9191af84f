"""bench.py's reference arm runs without a GPU (it times the unmodified reference on the host cores): check the JSON contract here."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line(reference):
    result = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", "--clips", "24"],
                            capture_output=True, text=True, check=True, cwd=ROOT, timeout=600)
    lines = [line for line in result.stdout.splitlines() if line.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "bone_poses_per_sec" and d["unit"] == "bone-poses/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_bench_refuses_to_run_the_product_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    result = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3", "--clips", "4"],
                            capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert result.returncode != 0
    assert "no CPU fallback" in (result.stderr + result.stdout)
