"""Randomised parity of the band SLIC kernel on the CPU emulator against the literal restatement: python tools/fuzz_emu_slic.py SECONDS [SEED]
(random shapes, steps, nc; flat frames, half-flat frames, frames with few colours -- exact ties).  163 batches in 400 s, seed 7: all equal."""
import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from depth_completion_mt_b200 import _lib, api, synth
from oracle import c_oracle as co
from tests.emu import build_emu
lib = _lib.bind(build_emu.build())
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
t0 = time.time(); n = 0
while time.time() - t0 < float(sys.argv[1]):
    rows = int(rng.integers(6, 80)); cols = int(rng.integers(10, 150)); step = int(rng.integers(4, 24)); nc = int(rng.integers(1, 80))
    frames = int(rng.integers(2, 4))
    labs = np.stack([synth.lab_image(int(rng.integers(0, 1000)), rows, cols) for _ in range(frames)])
    mode = rng.integers(0, 4)
    if mode == 1: labs[0] = int(rng.integers(0, 255))
    if mode == 2: labs[0, :, : cols // 2] = 17
    if mode == 3: labs = (labs // 64) * 64  # few colours: many exact ties
    bl, bc = api.generate_superpixels(labs, step, nc, return_centers=True, lib=lib)
    for f in range(frames):
        rl, rc = co.slic(labs[f], step, nc)
        ok = np.array_equal(bl[f], rl) and np.array_equal(np.isnan(bc[f]), np.isnan(rc)) and np.array_equal(bc[f][~np.isnan(rc)], rc[~np.isnan(rc)])
        if not ok:
            print("MISMATCH", rows, cols, step, nc, frames, mode, f, int((bl[f] != rl).sum())); sys.exit(1)
    n += 1
print("ok", n, "batches")
