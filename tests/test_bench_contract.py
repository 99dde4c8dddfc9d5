"""bench.py --impl reference: the JSON line the driver parses (runs the unmodified reference programs on the host cores)"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    if not (os.path.exists(os.path.join(ROOT, "oracle", "_ref", "encode")) and
            os.path.exists(os.path.join(ROOT, "oracle", "_ref", "decode"))):
        pytest.skip("oracle/_ref is not built (needs /root/reference)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "exactly one JSON line on stdout"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "encode/decode Mpixel/s, 8K RGB" and d["unit"] == "Mpixel/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert abs(d["cpu_baseline"]["value"] - d["value"]) < 1e-6 * max(1.0, d["value"])
    e = d["e2e"]
    assert e["unit"] == "Mpixel/s" and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert abs(e["value"] - d["value"]) < 1e-6 * max(1.0, d["value"])
    assert "workload" in d["config"] and "model" not in d["config"]
