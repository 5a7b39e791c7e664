"""bench.py --impl reference (the CPU arm the driver runs next to the GPU arm): one JSON line with the contract's keys for
every workload, at a small shape so that the whole file takes seconds.  The GPU arm needs a device and is exercised on the
box (profiles/r02_final_bench*.json)."""
from __future__ import annotations

import json
import os
import subprocess
import sys

import pytest

from tests.conftest import ROOT


@pytest.mark.parametrize("workload", ["lidar_only", "guided", "stereo", "stereo_chain", "lidar_camera_chain"])
def test_reference_arm_prints_the_contract_line(workload):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", workload, "--rows", "64", "--cols", "96",
           "--steps", "1", "--warmup", "1", "--frames", "8"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout  # only the JSON line reaches stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["steps"] == 1 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["gpu_launches"] == 0
    assert d["config"]["rows"] == 64 and d["config"]["cols"] == 96 and d["config"]["frames_per_gpu_per_step"] == 8
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["single_thread_ms"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
