"""bench.py's reference arm on this (GPU-less) machine: it must print exactly
ONE JSON line on stdout with the contract's keys, and never touch the GPU
package.  (The GPU arm is exercised by the driver on a B200.)"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(600)
def test_reference_arm_prints_one_json_line():
    env = dict(os.environ)
    env.pop("RANK", None), env.pop("WORLD_SIZE", None)
    p = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=580)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout[-1000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "icp_registrations_per_sec"
    assert d["unit"] == "registrations/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["config"]["workload"] == "c2_kitti64_120k_scan_to_scan_odometry"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "registrations/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == ""
