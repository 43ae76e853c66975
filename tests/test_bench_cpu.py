"""bench.py contract checks that need no GPU: the reference arm (the oracle port on the host cores) prints one JSON line with
the keys the driver reads, and the helper maths of the GPU arm."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "exactly one line on stdout"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["value"] > 0 and d["unit"] == "Gvoxel-updates/s"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["steps"] == 1 and d["n_gpus"] == 1
    # the other ranks of a torchrun launch exit without work
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True,
                       timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_algorithmic_bytes_match_the_survey_figure():
    """dense integration at 512^3, N = 4, 128x128x256 inverse volumes = 819.2 MB per frame (6.10 B per voxel-update)."""
    sys.path.insert(0, ROOT)
    import bench
    b = bench.algorithmic_bytes(512, 0, 0, 0, bricks=False)
    assert abs(b - (4 * 512 ** 3 + 16 * 128 * 128 * 256 * 4 + 16 * 512 * 424 * 4)) == 0
    assert abs(b / 1e6 - 819.2) < 0.1 and abs(b / 512 ** 3 - 6.10) < 0.01
    assert bench.peaks()[0] > 1000.0
    t, src = bench.measured_traffic(True)
    assert t is None or (t > 1e8 and "ncu" in src)
