"""bench.py prints exactly ONE JSON line on stdout, whatever native libraries write to fd 1 (NCCL's
version banner did land there under torchrun), and the CPU reference arm carries the keys the
driver reads. No GPU needed: the reference arm runs the reference's own loop (oracle/_ref) on the
host cores at a small frame."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_result_line_survives_stdout_chatter():
    code = ("import os, sys; sys.path.insert(0, %r); import bench; bench.claim_stdout(); "
            "os.write(1, b'NCCL version 2.28.9+cuda12.9\\n'); print('python chatter'); bench.emit({'a': 1})" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout == '{"a": 1}\n'
    assert "NCCL version" in r.stderr and "python chatter" in r.stderr


def test_reference_arm_prints_one_json_line(O):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--size", "96"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mpaths/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]
