"""The oracle itself: plain-C restatement (oracle/dcmt_oracle.c) against
  (1) the committed golden vectors (generated through OpenCV 4.13 by oracle/make_golden.py), and
  (2) when cv2 is importable, the cv2 transliteration on fresh seeds (SURVEY.md Appendix D recipes).
CPU only."""
from __future__ import annotations

import hashlib
import os

import numpy as np
import pytest

from depth_completion_mt_b200 import synth
from oracle import c_oracle as co
from oracle import cv2_oracle as cvo
from tests.conftest import GOLDEN, assert_bit_equal

needs_cv2 = pytest.mark.skipif(not cvo.HAVE_CV2, reason="cv2 not importable")

# float tolerances: Gaussian on non-q8 input, bilateral
GAUSS_TOL = 1e-4
BILATERAL_TOL = 2e-4


def lidar_case_names(g):
    return sorted({k.split("__")[0] for k in g.files})


def test_golden_operators(golden):
    g = golden["operators"]
    x = g["x"]
    assert_bit_equal(co.op_two_tap(x), g["two_tap"], "2-tap dilate")
    for k in (5, 7, 31):
        assert_bit_equal(co.op_box(x, k, 1), g[f"dilate{k}"], f"dilate{k}")
        assert_bit_equal(co.op_box(x, k, 0), g[f"erode{k}"], f"erode{k}")
    assert_bit_equal(co.op_box(co.op_box(x, 5, 1), 5, 0), g["close5"], "close5")
    assert_bit_equal(co.op_median5(x), g["median5"], "median5")
    assert_bit_equal(co.op_gaussian5(g["xq"]), g["gaussian5_q8"], "gaussian5 on q8 input is exact")
    assert np.abs(co.op_gaussian5(x) - g["gaussian5"]).max() <= GAUSS_TOL
    assert np.abs(co.op_bilateral5(x) - g["bilateral5"]).max() <= BILATERAL_TOL


def test_two_tap_quirk():
    """the int[5][5] diamond read as bytes has taps (-1,+1) and (+2,+2) only; pixels
    whose taps are both outside become -FLT_MAX: the whole last column plus (0, W-2) = H+1 pixels."""
    x = np.zeros((8, 10), np.float32)
    x[5, 5] = 7
    y = co.op_two_tap(x)
    assert y[3, 3] == 7 and y[6, 4] == 7 and (y == 7).sum() == 2
    neg = y == -np.finfo(np.float32).max
    assert neg.sum() == 8 + 1 and neg[:, -1].all() and neg[0, -2]


def test_golden_lidar_only(golden):
    g = golden["lidar_only"]
    for name in lidar_case_names(g):
        s = g[name + "__in"]
        q8 = name.startswith("q8") or name.startswith("multipass")
        st = {}
        got = co.img_completion(s, "none", st)
        assert_bit_equal(got, g[name + "__none"], f"{name} none")
        assert st["loop_passes"] == int(g[name + "__passes"]), name
        got = co.img_completion(s, "gaussian")
        if q8:
            assert_bit_equal(got, g[name + "__gaussian"], f"{name} gaussian (q8: exact)")
        else:
            assert np.abs(got - g[name + "__gaussian"]).max() <= GAUSS_TOL, name
        assert np.abs(co.img_completion(s, "bilateral") - g[name + "__bilateral"]).max() <= BILATERAL_TOL, name


def test_golden_multipass_needs_four_passes(golden):
    assert int(golden["lidar_only"]["multipass_120x64__passes"]) == 4


def test_golden_guided(golden):
    g = golden["guided"]
    for name in sorted({k.split("__")[0] for k in g.files}):
        s, lab, k = g[name + "__in"], g[name + "__labels"], int(g[name + "__k"])
        assert_bit_equal(co.interpolate_with_superpixels(s, lab, k, literal=True), g[name + "__sp1"], f"{name} literal loop")
        assert_bit_equal(co.interpolate_with_superpixels(s, lab, k, literal=False), g[name + "__sp1"], f"{name} closed form")
        assert_bit_equal(co.interpolate_with_superpixels(s, lab, k, use_superpixel=0), g[name + "__sp0"], f"{name} sp=0")
        assert_bit_equal(g[name + "__sp0"], co.img_completion(s, "gaussian"), f"{name}: use_superpixel=0 == img_completion")


def test_golden_stereo(golden):
    g = golden["stereo"]
    for name in sorted({k.split("__")[0] for k in g.files}):
        dig, left, right = g[name + "__depth_ig"], g[name + "__left"], g[name + "__right"]
        out, disp = co.stereo_refine(dig, left, right, final_gauss=False, return_disp=True)
        assert_bit_equal(out, g[name + "__default_nogauss"], f"{name} depth")
        assert_bit_equal(disp, g[name + "__disp4"], f"{name} disparity")
        assert np.abs(co.stereo_refine(dig, left, right) - g[name + "__default"]).max() <= GAUSS_TOL
        off = co.stereo_refine(dig, left, right, num_iterations=10, damp_factor=1370.0, err_clip=221.0, depth_clip=80.0,
                               final_gauss=False)
        assert_bit_equal(off, g[name + "__official10"], f"{name} OFFICIAL constants")
        dx, dy = co.measurement_derivatives(right.astype(np.float32))
        assert_bit_equal(dx, g[name + "__dx_right"], "dx")
        assert_bit_equal(dy, g[name + "__dy_right"], "dy")


def test_golden_full_size_digest():
    """352 x 1216 frames: the C oracle reproduces the sha256 of OpenCV's output (gaussian: q8 exact)."""
    lines = [l.split() for l in open(os.path.join(GOLDEN, "lidar_only_352x1216.sha256")) if not l.startswith("#")]
    for frame, kitti_like, blur, h_in, h_out in lines[:4]:
        s = synth.sparse_depth(int(frame), density=0.05, kitti_like=bool(int(kitti_like)))
        assert hashlib.sha256(s.tobytes()).hexdigest() == h_in, "synthetic input generator drifted"
        assert hashlib.sha256(co.img_completion(s, blur).tobytes()).hexdigest() == h_out, (frame, kitti_like, blur)


def test_column_extrapolation_cases():
    """Appendix A5: empty column -> all 100; single valid pixel -> whole column that value."""
    d = np.zeros((6, 3), np.float32)
    d[2, 1] = 5.0
    d[1, 2] = 7.0
    d[4, 2] = 9.0
    o = co.op_column_extrapolation(d)
    assert (o[:, 0] == 100.0).all() and (o[:, 1] == 5.0).all()
    assert list(o[:, 2]) == [7.0, 7.0, 0.0, 0.0, 9.0, 9.0]


def test_threshold_is_float_0p1():
    """`d > 0.1` (double literal) <=> d >= 0.1f."""
    t = np.float32(0.1)
    below = np.nextafter(t, np.float32(0))
    s = np.array([[t, below]], np.float32)
    o = co.img_completion(s, "none")
    assert np.isfinite(o).all()
    for v, valid in ((t, True), (below, False)):
        assert (float(v) > 0.1) == valid


@needs_cv2
@pytest.mark.parametrize("shape,density", [((352, 1216), 0.05), ((352, 1216), 0.01), ((97, 211), 0.03), ((31, 17), 0.2), ((4, 300), 0.1)])
def test_c_oracle_equals_cv2_lidar(shape, density):
    for f in range(2):
        s = synth.sparse_depth(20 + f, shape[0], shape[1], density, kitti_like=bool(f))
        for bt in ("gaussian", "none"):
            a, b = {}, {}
            assert_bit_equal(co.img_completion(s, bt, a), cvo.img_completion(s, bt, b), f"{shape} {bt}")
            assert a["loop_passes"] == b["loop_passes"]
    s = synth.sparse_depth_float(5, shape[0], shape[1], density)
    assert_bit_equal(co.img_completion(s, "none"), cvo.img_completion(s, "none"), "float input, no blur")
    assert np.abs(co.img_completion(s, "gaussian") - cvo.img_completion(s, "gaussian")).max() <= GAUSS_TOL
    assert np.abs(co.img_completion(s, "bilateral") - cvo.img_completion(s, "bilateral")).max() <= BILATERAL_TOL


@needs_cv2
def test_c_oracle_equals_cv2_guided_and_stereo():
    s = synth.sparse_depth(9, 80, 120, 0.05)
    lab, k = synth.superpixel_labels(9, 80, 120, 12)
    want = cvo.interpolate_with_superpixels(s, lab, k)
    assert_bit_equal(co.interpolate_with_superpixels(s, lab, k, literal=True), want, "guided literal")
    assert_bit_equal(co.interpolate_with_superpixels(s, lab, k, literal=False), want, "guided closed form")
    dig, left, right = synth.stereo_pair(9, 60, 180)
    assert_bit_equal(co.stereo_refine(dig, left, right, final_gauss=False), cvo.stereo_refine(dig, left, right, final_gauss=False), "stereo")


@needs_cv2
def test_bilateral_in_place_throws_in_opencv():
    """the reference's in-place bilateralFilter call asserts inside OpenCV."""
    import cv2

    y = np.random.default_rng(0).random((20, 20)).astype(np.float32)
    with pytest.raises(cv2.error):
        cv2.bilateralFilter(y, 5, 1.5, 2.0, dst=y)


@needs_cv2
def test_golden_files_are_reproducible(tmp_path, monkeypatch):
    """oracle/make_golden.py regenerates the committed lidar fixture bit for bit."""
    from oracle import make_golden

    monkeypatch.setattr(make_golden, "OUT", str(tmp_path))
    make_golden.main()
    for name in ("lidar_only", "guided", "stereo", "operators"):
        a, b = np.load(os.path.join(GOLDEN, name + ".npz")), np.load(os.path.join(str(tmp_path), name + ".npz"))
        assert sorted(a.files) == sorted(b.files)
        for k in a.files:
            assert_bit_equal(np.asarray(a[k]), np.asarray(b[k]), f"{name}:{k}")
    assert open(os.path.join(GOLDEN, "lidar_only_352x1216.sha256")).read() == open(os.path.join(str(tmp_path), "lidar_only_352x1216.sha256")).read()
