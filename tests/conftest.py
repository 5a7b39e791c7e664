"""Shared fixtures.

Backends under test
  * ``gpu``  -- the product: depth_completion_mt_b200/libdcmt.so on cuda:0, through the C ABI
                (tests carrying @pytest.mark.gpu; they FAIL, not skip, when the library is missing).
  * ``emu``  -- the same kernel sources compiled with g++ against tests/emu/cuda_emu.h and executed on
                the CPU by a fiber emulator of the CUDA execution model (test infrastructure; lets the
                CPU-only suite diff real kernel logic against the oracle).
The checker is always the oracle (oracle/), pinned by tests/golden/.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def emu_lib():
    from depth_completion_mt_b200 import _lib
    from tests.emu import build_emu

    return _lib.bind(build_emu.build())


@pytest.fixture(scope="session")
def gpu_lib():
    """The product library on a real device.  No skip: a missing library or device is a failure."""
    import torch

    from depth_completion_mt_b200 import _lib

    assert torch.cuda.is_available(), "gpu-marked test needs a CUDA device"
    lib = _lib.load()
    assert lib.dcmt_device_count() >= 1
    return lib


@pytest.fixture(scope="session")
def golden():
    return {name: np.load(os.path.join(GOLDEN, name + ".npz")) for name in ("lidar_only", "operators", "guided", "stereo")}


def assert_bit_equal(got: np.ndarray, want: np.ndarray, what: str = ""):
    got = np.ascontiguousarray(got)
    want = np.ascontiguousarray(want)
    assert got.shape == want.shape and got.dtype == want.dtype, f"{what}: shape/dtype {got.shape}/{got.dtype} vs {want.shape}/{want.dtype}"
    if got.dtype.kind == "f":
        a, b = got.view(np.uint32), want.view(np.uint32)
    else:
        a, b = got, want
    if not np.array_equal(a, b):
        bad = np.argwhere(a != b)
        i = tuple(bad[0])
        raise AssertionError(f"{what}: {len(bad)} of {got.size} elements differ; first at {i}: got {got[i]!r} want {want[i]!r}")
