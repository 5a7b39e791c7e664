"""Randomised parity run of the fused and generic pipelines on the CPU emulator against the C oracle (test
infrastructure; not part of the pytest suite because it runs for as long as asked):
    python tools/fuzz_emu.py [seconds] [seed]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from depth_completion_mt_b200 import _lib, api, synth  # noqa: E402
from oracle import c_oracle as co  # noqa: E402
from tests.emu import build_emu  # noqa: E402

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 60
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rng = np.random.default_rng(seed)
lib = _lib.bind(build_emu.build())
t_end = time.time() + secs
n = 0
while time.time() < t_end:
    rows = int(rng.integers(32, 230))
    cols = int(rng.integers(32, 400))
    p = float(rng.choice([0.001, 0.003, 0.01, 0.03, 0.05, 0.1, 0.2, 0.5]))
    kitti = bool(rng.integers(0, 2))
    blur = str(rng.choice(["gaussian", "none"]))
    u16 = bool(rng.integers(0, 2))
    d16 = synth.sparse_depth_q8(int(rng.integers(0, 1 << 30)), rows, cols, p, kitti_like=kitti)
    if rng.random() < 0.3:  # sprinkle boundary codes
        ys, xs = rng.integers(0, rows, 20), rng.integers(0, cols, 20)
        d16[ys, xs] = rng.choice([1, 25, 26, 27, 25574, 25575, 25600, 30000, 65535], 20)
    s = d16.astype(np.float32) / np.float32(256)
    st = {}
    want = co.img_completion(s, blur, st)
    got, stats = api.img_completion(d16 if u16 else s, False, blur, return_stats=True, lib=lib)
    ok = np.array_equal(got.view(np.uint32), want.view(np.uint32)) and int(stats[0, 0]) == st["loop_passes"] and int(stats[0, 1]) == st["holes_before_loop"]
    if not ok:
        print(f"MISMATCH rows={rows} cols={cols} p={p} kitti={kitti} blur={blur} u16={u16} diff={(got != want).sum()} stats={stats[0]} ref={st}", flush=True)
        np.save(f"/tmp/fuzz_fail_{n}.npy", d16)
        sys.exit(1)
    n += 1
print(f"fuzz ok: {n} random frames, seed {seed}")
