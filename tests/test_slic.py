"""SLIC superpixels against the literal restatement of Slic::generate_superpixels in the C oracle.
Bar: labels identical for every pixel, centres bit-equal (including NaN centres of empty clusters); the labels feed
interpolate_with_superpixels like main_lc.cpp:201-220."""
from __future__ import annotations

import numpy as np
import pytest

from depth_completion_mt_b200 import api, synth
from oracle import c_oracle as co


def body(lib, to_backend, cases, big_batch=None):
    for k, (rows, cols, step, nc) in enumerate(cases):
        lab = synth.lab_image(k, rows, cols)
        if k == 1:  # flat image: every distance ties on colour, exercises "lowest centre index wins"
            lab[:] = 77
        labels, centers = api.generate_superpixels(to_backend(lab), step, nc, return_centers=True, lib=lib)
        labels, centers = (a if isinstance(a, np.ndarray) else a.cpu().numpy() for a in (labels, centers))
        ref_l, ref_c = co.slic(lab, step, nc)
        assert centers.shape == ref_c.shape == (lib.dcmt_slic_center_count(rows, cols, int(step)), 5)
        assert np.array_equal(labels, ref_l), f"case {k}: {(labels != ref_l).sum()} of {labels.size} labels differ"
        assert np.array_equal(centers.view(np.uint64), ref_c.view(np.uint64)) or np.array_equal(np.isnan(centers), np.isnan(ref_c)) and \
            np.array_equal(centers[~np.isnan(centers)], ref_c[~np.isnan(ref_c)]), f"case {k}: centres differ"
    # the labels drive the guided completion (main_lc.cpp:201-220)
    rows, cols = 96, 160
    lab = synth.lab_image(9, rows, cols)
    labels = api.generate_superpixels(to_backend(lab), 12.7, 40, lib=lib)
    ln = labels if isinstance(labels, np.ndarray) else labels.cpu().numpy()
    kc = lib.dcmt_slic_center_count(rows, cols, 12)
    sparse = synth.sparse_depth(9, rows, cols, 0.08)
    out = api.interpolate_with_superpixels(labels, to_backend(sparse), "gaussian", 1, n_clusters=kc, lib=lib)
    on = out if isinstance(out, np.ndarray) else out.cpu().numpy()
    ref = co.interpolate_with_superpixels(sparse, co.slic(lab, 12, 40)[0], kc)
    assert np.array_equal(ln, co.slic(lab, 12, 40)[0])
    assert np.array_equal(on.view(np.uint32), ref.view(np.uint32))
    # a batch of frames runs side by side and gives the per-frame results
    labs = np.stack([synth.lab_image(20 + f, 64, 96) for f in range(3)])
    bl, bc = api.generate_superpixels(to_backend(labs), 10, 40, return_centers=True, lib=lib)
    bl, bc = (a if isinstance(a, np.ndarray) else a.cpu().numpy() for a in (bl, bc))
    for f in range(3):
        rl, rc = co.slic(labs[f], 10, 40)
        assert np.array_equal(bl[f], rl), f"batch frame {f}"
        assert np.array_equal(bc[f].view(np.uint64), rc.view(np.uint64)), f"batch centres {f}"
    # a larger batch (bands of whole frames in k_slic_assign_band): same labels and centres, incl. a flat frame (all ties: every
    # pixel goes through slic_resolve) and an odd number of centres
    shapes = big_batch or (36, 52, 8)
    labs = np.stack([synth.lab_image(40 + f, shapes[0], shapes[1]) for f in range(4)] * 8)
    labs[5] = 77
    bl, bc = api.generate_superpixels(to_backend(labs), shapes[2], 40, return_centers=True, lib=lib)
    bl, bc = (a if isinstance(a, np.ndarray) else a.cpu().numpy() for a in (bl, bc))
    for f in (0, 1, 2, 3, 5, 31):
        rl, rc = co.slic(labs[f], shapes[2], 40)
        assert np.array_equal(bl[f], rl), f"band batch frame {f}: {(bl[f] != rl).sum()} labels differ"
        assert np.array_equal(np.isnan(bc[f]), np.isnan(rc)) and np.array_equal(bc[f][~np.isnan(rc)], rc[~np.isnan(rc)]), f"band batch centres {f}"
    # argument errors
    with pytest.raises(Exception):
        api.generate_superpixels(to_backend(lab), 3, 40, lib=lib)


def test_emu_slic(emu_lib):
    body(emu_lib, lambda a: a, [(64, 96, 10, 40), (50, 70, 9, 30), (80, 120, 18, 50)])


def check_tile_kernel(lib, to_backend, rows, cols, step):
    """More centres than the band kernel's shared memory holds (K x 72 bytes > 200 KB): the 16 x 16 tile kernel runs."""
    k = lib.dcmt_slic_center_count(rows, cols, step)
    assert k * 72 > 200 * 1024
    lab = synth.lab_image(61, rows, cols)
    labels, centers = api.generate_superpixels(to_backend(lab), step, 40, return_centers=True, lib=lib)
    labels, centers = (a if isinstance(a, np.ndarray) else a.cpu().numpy() for a in (labels, centers))
    ref_l, ref_c = co.slic(lab, step, 40)
    assert np.array_equal(labels, ref_l), f"{(labels != ref_l).sum()} of {labels.size} labels differ"
    assert np.array_equal(np.isnan(centers), np.isnan(ref_c)) and np.array_equal(centers[~np.isnan(ref_c)], ref_c[~np.isnan(ref_c)])


@pytest.mark.parametrize("rows,cols,step,nc,frames", [
    (37, 70, 7, 20, 2),     # partial last strip (70 = 2 x 32 + 6), bands of a few rows
    (12, 200, 5, 40, 3),    # low and wide: more bands than bins in y
    (64, 33, 9, 10, 2),     # one pixel in the second strip
    (90, 130, 31, 60, 2),   # step about as large as a strip: few, wide windows
    (41, 97, 4, 5, 2),      # smallest step, colour dominates (nc small): many near ties on a noisy image
    (6, 40, 4, 30, 2),      # a single row of centres
])
def test_emu_slic_band_kernel_shapes(emu_lib, rows, cols, step, nc, frames):
    """The band kernel (batches) at shapes that stress its item walk: partial strips, bands of very few rows, steps from 4 to
    strip width.  One frame of each batch is flat (every pixel deferred to slic_resolve), one has a flat half (long runs)."""
    check_band_shapes(emu_lib, lambda a: a, rows, cols, step, nc, frames)


def check_band_shapes(lib, to_backend, rows, cols, step, nc, frames):
    labs = np.stack([synth.lab_image(90 + f, rows, cols) for f in range(frames)])
    labs[0, :, cols // 2:] = 40
    if frames > 2:
        labs[2] = 200
    bl, bc = api.generate_superpixels(to_backend(labs), step, nc, return_centers=True, lib=lib)
    bl, bc = (a if isinstance(a, np.ndarray) else a.cpu().numpy() for a in (bl, bc))
    for f in range(frames):
        rl, rc = co.slic(labs[f], step, nc)
        assert np.array_equal(bl[f], rl), f"frame {f}: {(bl[f] != rl).sum()} of {rl.size} labels differ"
        assert np.array_equal(np.isnan(bc[f]), np.isnan(rc)) and np.array_equal(bc[f][~np.isnan(rc)], rc[~np.isnan(rc)]), f"centres {f}"


def test_emu_slic_band_kernel_without_candidate_list():
    """More centres around a strip than a warp has lanes make the band kernel hand every pixel of those rows to slic_resolve (old
    labels kept under the marker, incl. pixels no window covers).  DCMT_SLIC_CAND_MAX lowers that limit so that ordinary frames take
    the path -- 0: all rows, 6: some rows; a single frame through the band kernel as well (separate process: read once)."""
    import os
    import subprocess
    import sys

    from tests.conftest import ROOT

    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import numpy as np\n"
        "from depth_completion_mt_b200 import _lib, api, synth\n"
        "from oracle import c_oracle as co\n"
        "from tests.emu import build_emu\n"
        "lib = _lib.bind(build_emu.build())\n"
        "lab = np.stack([synth.lab_image(70 + k, 50, 88) for k in range(2)])\n"
        "lab[1, :, 40:] = 9\n"
        "labels, centers = api.generate_superpixels(lab, 9, 30, return_centers=True, lib=lib)\n"
        "for k in range(2):\n"
        "    rl, rc = co.slic(lab[k], 9, 30)\n"
        "    assert np.array_equal(labels[k], rl), (k, int((labels[k] != rl).sum()))\n"
        "    assert np.array_equal(np.isnan(centers[k]), np.isnan(rc)) and np.array_equal(centers[k][~np.isnan(rc)], rc[~np.isnan(rc)]), k\n"
        "print('CAND_OK')\n" % ROOT
    )
    for cand_max in ("0", "6"):
        env = dict(os.environ, DCMT_SLIC_CAND_MAX=cand_max, DCMT_SLIC_BAND_MIN_FRAMES="1")
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900, env=env)
        assert r.returncode == 0 and "CAND_OK" in r.stdout, r.stdout + r.stderr


def test_emu_slic_tile_kernel_on_a_batch():
    """Batches take the band kernel unless the frame's centres exceed its shared memory (the GPU test below does that at 352 x 1216,
    step 10); here DCMT_SLIC_BAND_MIN_FRAMES keeps a small batch on the 16 x 16 tile kernel (separate process: read once)."""
    import os
    import subprocess
    import sys

    from tests.conftest import ROOT

    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import numpy as np\n"
        "from depth_completion_mt_b200 import _lib, api, synth\n"
        "from oracle import c_oracle as co\n"
        "from tests.emu import build_emu\n"
        "lib = _lib.bind(build_emu.build())\n"
        "lab = np.stack([synth.lab_image(80 + k, 64, 96) for k in range(3)])\n"
        "lab[2] = 5\n"
        "labels, centers = api.generate_superpixels(lab, 10, 40, return_centers=True, lib=lib)\n"
        "for k in range(3):\n"
        "    rl, rc = co.slic(lab[k], 10, 40)\n"
        "    assert np.array_equal(labels[k], rl), (k, int((labels[k] != rl).sum()))\n"
        "    assert np.array_equal(np.isnan(centers[k]), np.isnan(rc)) and np.array_equal(centers[k][~np.isnan(rc)], rc[~np.isnan(rc)]), k\n"
        "print('TILE_OK')\n" % ROOT
    )
    env = dict(os.environ, DCMT_SLIC_BAND_MIN_FRAMES="1000")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0 and "TILE_OK" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["host", "device"])
def test_gpu_slic(gpu_lib, mode):
    import torch

    body(gpu_lib, (lambda a: a) if mode == "host" else (lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()),
         [(64, 96, 10, 40), (50, 70, 9, 30), (352, 1216, 18, 50), (375, 1242, 68, 40)], big_batch=(352, 1216, 18))


@pytest.mark.gpu
def test_gpu_slic_tile_kernel(gpu_lib):
    import torch

    check_tile_kernel(gpu_lib, lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda(), 352, 1216, 10)


@pytest.mark.gpu
def test_gpu_slic_band_kernel_shapes(gpu_lib):
    """The emulator's stress shapes on the device, plus KITTI raw size (1242 = 38 x 32 + 26 columns, 375 rows) in a small batch."""
    import torch

    to_dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    for case in [(37, 70, 7, 20, 2), (12, 200, 5, 40, 3), (64, 33, 9, 10, 2), (90, 130, 31, 60, 2), (41, 97, 4, 5, 2), (6, 40, 4, 30, 2),
                 (375, 1242, 18, 50, 3)]:
        check_band_shapes(gpu_lib, to_dev, *case)
