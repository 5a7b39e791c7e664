"""(e) multi-GPU plumbing on CPU: frame sharding covers the batch exactly once, and the validation
gather works over gloo with world_size 2 (the N>1 path of bench.py, minus the GPU)."""
from __future__ import annotations

import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from depth_completion_mt_b200 import sharding
from tests.conftest import ROOT


def test_shard_ranges_partition_the_batch():
    for n in (0, 1, 7, 64, 1024, 1025):
        for world in (1, 2, 3, 4, 8):
            cover = []
            for r in range(world):
                lo, hi = sharding.shard_range(n, r, world)
                assert 0 <= lo <= hi <= n
                cover += list(range(lo, hi))
            assert cover == list(range(n))
            sizes = [sharding.shard_range(n, r, world)[1] - sharding.shard_range(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(8, 2, 2)


def test_frame_checksums_detect_single_bit_flips():
    a = np.random.default_rng(0).random((3, 16, 20)).astype(np.float32)
    b = a.copy()
    b[1, 5, 7] = np.nextafter(b[1, 5, 7], np.float32(2))
    ca, cb = sharding.frame_checksums(a), sharding.frame_checksums(b)
    assert ca[0] == cb[0] and ca[2] == cb[2] and ca[1] != cb[1]


WORKER = r"""
import os, sys, json
sys.path.insert(0, {root!r})
import numpy as np, torch
from depth_completion_mt_b200 import sharding, synth
rank, local_rank, world = sharding.init_process_group("gloo")
n = 10
lo, hi = sharding.shard_range(n, rank, world)
frames = np.stack([synth.sparse_depth(f, 12, 20, 0.2) for f in range(lo, hi)])
res = sharding.gather_validation(10.0 + rank, hi - lo, sharding.frame_checksums(frames), device=torch.device("cpu"))
allsums = torch.cat(res["checksums"])
want = sharding.frame_checksums(np.stack([synth.sparse_depth(f, 12, 20, 0.2) for f in range(n)]))
assert torch.equal(allsums, want), (allsums, want)
assert res["max_ms"] == 10.0 + world - 1 and res["total_frames"] == n and res["world"] == world
if rank == 0:
    print("GATHER_OK", json.dumps(res["frames_per_rank"]))
"""


def test_gloo_world_size_2_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "GATHER_OK [5, 5]" in r.stdout
