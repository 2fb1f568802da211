"""CPU: the reference arm of bench.py (`--impl reference`: the unmodified lib_icer of oracle/_ref on the host cores) runs
without a GPU and prints the JSON line the driver expects (same metric / unit / config as the GPU arm)."""
import json
import os
import subprocess
import sys

import pytest

import harness

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not harness.have_ref(), reason="oracle/_ref not built")
def test_reference_arm_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=240, cwd=ROOT, env=dict(os.environ, ICER_BENCH_FRAMES_PER_CORE="3"))
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "encode+decode Mpixels/s" and line["unit"] == "Mpixels/s"
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["higher_is_better"] is True
    for k in ("n_gpus", "steps", "warmup", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert "workload" in line["config"] and "512x512" in line["config"]["workload"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    e = line["e2e"]
    assert e["value"] == line["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
