"""Property-based parity (hypothesis) of the fused strict-q8 path on the CPU emulator: random shapes, densities, boundary
codes, uint16 / float32 input, guided labels -- always the oracle's bytes, always the oracle's loop statistics."""
from __future__ import annotations

import numpy as np
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from depth_completion_mt_b200 import api
from oracle import c_oracle as co

BOUNDARY = [1, 25, 26, 27, 300, 25573, 25574, 25575, 25600, 30000, 65535]


@st.composite
def frames(draw):
    rows = draw(st.integers(32, 110))
    cols = draw(st.integers(32, 180))
    density = draw(st.sampled_from([0.002, 0.01, 0.05, 0.2, 0.6]))
    seed = draw(st.integers(0, 2**31 - 1))
    rng = np.random.default_rng(seed)
    d16 = np.where(rng.random((rows, cols)) < density, rng.integers(26, 25575, (rows, cols)), 0).astype(np.uint16)
    if draw(st.booleans()):
        n = draw(st.integers(1, 40))
        d16[rng.integers(0, rows, n), rng.integers(0, cols, n)] = rng.choice(BOUNDARY, n)
    if draw(st.booleans()):  # an empty band: column extrapolation and multi-pass fills
        r0 = draw(st.integers(0, rows - 1))
        d16[r0: r0 + draw(st.integers(1, rows))] = 0
    return d16


@settings(max_examples=25, deadline=None, suppress_health_check=list(HealthCheck))
@given(d16=frames(), blur=st.sampled_from(["gaussian", "none"]), as_u16=st.booleans())
def test_emu_fused_matches_oracle(emu_lib, d16, blur, as_u16):
    s = d16.astype(np.float32) / np.float32(256)
    ref_st = {}
    want = co.img_completion(s, blur, ref_st)
    got, stats = api.img_completion(d16 if as_u16 else s, False, blur, return_stats=True, lib=emu_lib)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert int(stats[0, 0]) == ref_st["loop_passes"] and int(stats[0, 1]) == ref_st["holes_before_loop"]
    assert int(stats[0, 2]) == ref_st["holes_after_extrapolation"] and int(stats[0, 3]) == 1


@settings(max_examples=12, deadline=None, suppress_health_check=list(HealthCheck))
@given(d16=frames(), step=st.integers(5, 24), seed=st.integers(0, 2**31 - 1))
def test_emu_fused_guided_matches_oracle(emu_lib, d16, step, seed):
    rows, cols = d16.shape
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:rows, 0:cols]
    per_row = (cols + step + 2) // step + 1
    lab = (((yy + rng.integers(-2, 3, (rows, cols))).clip(0) // step) * per_row + (xx + rng.integers(-2, 3, (rows, cols))).clip(0) // step).astype(np.int32)
    lab[rng.random((rows, cols)) < 0.03] = -1
    k = int(lab.max()) + 1
    s = d16.astype(np.float32) / np.float32(256)
    want = co.interpolate_with_superpixels(s, lab, k, literal=False)
    got, stats = api.interpolate_with_superpixels(lab, s, "gaussian", 1, n_clusters=k, return_stats=True, lib=emu_lib)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)) and int(stats[0, 3]) == 1
