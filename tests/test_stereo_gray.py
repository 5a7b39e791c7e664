"""cv::cvtColor(COLOR_BGR2GRAY) (main_sl.cpp:1167,1171), the step in front of the EntryType fill: bit-equal to cv2 4.13
(OpenCV's fixed-point 8-bit path) for every (B, G, R) on a grid, random images, odd widths and pitched rows."""
from __future__ import annotations

import numpy as np
import pytest

from depth_completion_mt_b200 import api

cv2 = pytest.importorskip("cv2")


def body(lib, to_backend):
    rng = np.random.default_rng(0)
    vals = np.array(list(range(0, 256, 5)) + [255], np.uint8)
    B, G, R = np.meshgrid(vals, vals, vals, indexing="ij")
    cube = np.stack([B, G, R], -1).reshape(len(vals), len(vals) * len(vals), 3)
    for img in (cube, rng.integers(0, 256, (97, 171, 3), dtype=np.uint8), rng.integers(0, 256, (5, 1, 3), dtype=np.uint8),
                rng.integers(0, 256, (3, 4, 33, 3), dtype=np.uint8)):
        got = api.bgr2gray(to_backend(np.ascontiguousarray(img)), lib=lib)
        got = got if isinstance(got, np.ndarray) else got.cpu().numpy()
        want = np.stack([cv2.cvtColor(f, cv2.COLOR_BGR2GRAY) for f in img]) if img.ndim == 4 else cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
        assert np.array_equal(got, want.reshape(got.shape)), img.shape


def test_emu_bgr2gray(emu_lib):
    body(emu_lib, lambda a: a)
    # pitched rows through the C ABI
    rng = np.random.default_rng(1)
    rows, cols, bp, gp = 9, 13, 47, 20
    src = rng.integers(0, 256, (rows, bp), dtype=np.uint8)
    dst = np.full((rows, gp), 7, np.uint8)
    emu_lib.check(emu_lib.dcmt_bgr2gray_u8_host(src.ctypes.data, dst.ctypes.data, rows, cols, bp, gp, 1))
    want = cv2.cvtColor(np.ascontiguousarray(src[:, : cols * 3]).reshape(rows, cols, 3), cv2.COLOR_BGR2GRAY)
    assert np.array_equal(dst[:, :cols], want) and (dst[:, cols:] == 7).all()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["host", "device"])
def test_gpu_bgr2gray(gpu_lib, mode):
    import torch

    body(gpu_lib, (lambda a: a) if mode == "host" else (lambda a: torch.from_numpy(a).cuda()))
