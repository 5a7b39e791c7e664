"""Pinning the oracles to THE REFERENCE ITSELF.

oracle/_ref/libdcmt_ref.so = the reference's img_completion.cpp, img_completion_lc.cpp and slic.cpp compiled unmodified
from /root/reference against the stand-in OpenCV header of oracle/refshim/ (containers only; dilate / morphologyEx /
medianBlur / GaussianBlur / bilateralFilter are forwarded to cv2 = OpenCV 4.13).  These tests run the reference's own
compiled code and require the restatements the GPU parity tests check against -- oracle/dcmt_oracle.c (plain C) and
oracle/cv2_oracle.py (transliteration) -- and the committed golden vectors to agree with it bit for bit.

CPU only.  In the build container the library is rebuilt from /root/reference on demand; on a box without the
reference sources the prebuilt file that travelled with the snapshot is used; with neither, the tests skip.
"""
from __future__ import annotations

import hashlib
import os

import numpy as np
import pytest

from depth_completion_mt_b200 import synth
from oracle import c_oracle as co
from oracle import cv2_oracle as cvo
from oracle import ref_oracle as ro
from tests.conftest import GOLDEN, assert_bit_equal

pytestmark = pytest.mark.skipif(not ro.available(), reason="reference build (oracle/_ref/libdcmt_ref.so) or cv2 not available")


def test_library_is_built_from_the_reference_tree():
    so = ro.build()
    assert os.path.basename(os.path.dirname(so)) == "_ref"
    if ro.have_sources():  # the recipe compiles the sources where they lie; nothing of the reference is copied into the repo
        mk = open(os.path.join(os.path.dirname(os.path.dirname(so)), "Makefile")).read()
        assert "$(REF)/src/DC_lidar_only/img_completion.cpp" in mk and "$(REF)/src/DC_lidar_camera/slic.cpp" in mk


@pytest.mark.parametrize("shape,density,kitti_like", [
    ((352, 1216), 0.05, False), ((352, 1216), 0.05, True), ((352, 1216), 0.01, False), ((97, 211), 0.03, False),
    ((120, 64), 0.002, False), ((31, 17), 0.2, False), ((4, 300), 0.1, False), ((1, 1), 1.0, False), ((5, 1), 0.5, False),
])
def test_img_completion_reference_vs_restatements(shape, density, kitti_like):
    """a1 (img_completion.cpp:17-204), q8 input: reference == C oracle == cv2 transliteration, every bit, none and gaussian."""
    s = synth.sparse_depth(40, shape[0], shape[1], density, kitti_like=kitti_like)
    for bt in ("none", "gaussian"):
        ref = ro.img_completion(s, bt)
        assert_bit_equal(co.img_completion(s, bt), ref, f"C oracle vs reference {shape} {bt}")
        assert_bit_equal(cvo.img_completion(s, bt), ref, f"cv2 transliteration vs reference {shape} {bt}")


def test_img_completion_reference_float_input():
    """Non-q8 float input: min/max/median/fill stages stay bit-exact (blur none); the float Gaussian within 1e-4."""
    s = synth.sparse_depth_float(3, 97, 211, 0.05)
    assert_bit_equal(co.img_completion(s, "none"), ro.img_completion(s, "none"), "float input, blur none")
    assert_bit_equal(cvo.img_completion(s, "gaussian"), ro.img_completion(s, "gaussian"), "transliteration uses the same OpenCV")
    assert np.abs(co.img_completion(s, "gaussian") - ro.img_completion(s, "gaussian")).max() <= 1e-4


def test_img_completion_reference_special_values():
    """Negatives, values that invert below 0.1, the exact threshold, an empty frame and an all-valid frame."""
    t = np.float32(0.1)
    s = np.zeros((40, 48), np.float32)
    s[3, 5], s[10, 10], s[11, 40], s[20, 7], s[30, 30] = -4.0, 99.95, t, np.nextafter(t, np.float32(0)), 100.0
    s[25, 12] = 12.5
    assert_bit_equal(co.img_completion(s, "none"), ro.img_completion(s, "none"), "special values none")
    assert np.abs(co.img_completion(s, "gaussian") - ro.img_completion(s, "gaussian")).max() <= 1e-4  # 99.95 etc. are not q8
    for frame in (np.zeros((33, 35), np.float32), np.full((33, 35), 7.25, np.float32)):
        for bt in ("none", "gaussian"):
            assert_bit_equal(co.img_completion(frame, bt), ro.img_completion(frame, bt), f"constant frame {bt}")


def test_reference_bilateral_branch_throws():
    """img_completion.cpp:174 calls cv::bilateralFilter in place; OpenCV asserts src.data != dst.data."""
    with pytest.raises(RuntimeError, match="src.data != dst.data"):
        ro.img_completion(synth.sparse_depth(1, 40, 60, 0.05), "bilateral")


def test_reference_reproduces_golden_vectors(golden):
    """The committed fixtures (made through cv2 by oracle/make_golden.py) are what the reference itself computes."""
    g = golden["lidar_only"]
    for name in sorted({k.split("__")[0] for k in g.files}):
        s = g[name + "__in"]
        assert_bit_equal(ro.img_completion(s, "none"), g[name + "__none"], f"{name} none")
        assert_bit_equal(ro.img_completion(s, "gaussian"), g[name + "__gaussian"], f"{name} gaussian")
    g = golden["guided"]
    for name in sorted({k.split("__")[0] for k in g.files}):
        s, lab, k = g[name + "__in"], g[name + "__labels"], int(g[name + "__k"])
        assert_bit_equal(ro.interpolate_with_superpixels(lab, s, use_superpixel=1, n_clusters=k), g[name + "__sp1"], f"{name} sp=1")
        assert_bit_equal(ro.interpolate_with_superpixels(lab, s, use_superpixel=0, n_clusters=k), g[name + "__sp0"], f"{name} sp=0")


def test_reference_reproduces_full_size_digests():
    lines = [l.split() for l in open(os.path.join(GOLDEN, "lidar_only_352x1216.sha256")) if not l.startswith("#")]
    assert lines
    for frame, kitti_like, blur, h_in, h_out in lines:
        s = synth.sparse_depth(int(frame), density=0.05, kitti_like=bool(int(kitti_like)))
        assert hashlib.sha256(s.tobytes()).hexdigest() == h_in
        assert hashlib.sha256(ro.img_completion(s, blur).tobytes()).hexdigest() == h_out, (frame, kitti_like, blur)


@pytest.mark.parametrize("shape,step", [((64, 96), 18), ((50, 70), 12)])
def test_interpolate_with_superpixels_reference_vs_restatements(shape, step):
    """a2 (img_completion_lc.cpp:34-203): the reference's per-superpixel loop == the literal C loop == the closed form the
    kernels implement, incl. unassigned pixels (-1) and labels >= slic.centers.size() (never visited)."""
    rows, cols = shape
    s = synth.sparse_depth(7, rows, cols, 0.08)
    lab, k = synth.superpixel_labels(7, rows, cols, step=step)
    lab = lab.copy()
    lab[5:9, 10:14] = -1
    for n_clusters in (k, k - 3):
        ref = ro.interpolate_with_superpixels(lab, s, use_superpixel=1, n_clusters=n_clusters)
        assert_bit_equal(co.interpolate_with_superpixels(s, lab, n_clusters, literal=True), ref, "literal C loop vs reference")
        assert_bit_equal(co.interpolate_with_superpixels(s, lab, n_clusters, literal=False), ref, "closed form vs reference")
        assert_bit_equal(cvo.interpolate_with_superpixels(s, lab, n_clusters), ref, "cv2 transliteration vs reference")
    ref0 = ro.interpolate_with_superpixels(lab, s, use_superpixel=0, n_clusters=k)
    assert_bit_equal(co.interpolate_with_superpixels(s, lab, k, use_superpixel=0), ref0, "use_superpixel = 0")
    assert_bit_equal(ref0, ro.img_completion(s, "gaussian"), "use_superpixel = 0 is img_completion with the gaussian blur")


@pytest.mark.parametrize("shape,step,nc", [((352, 1216), 18, 40), ((120, 200), 18, 40), ((64, 90), 10, 20), ((40, 40), 50, 40)])
def test_slic_reference_vs_restatement(shape, step, nc):
    """f1 (slic.cpp:19-182): labels identical for every pixel, centres bit-equal (NaN centres of empty clusters included)."""
    img = synth.lab_image(2, shape[0], shape[1])
    r_lab, r_cen, r_cnt = ro.generate_superpixels(img, step, nc)
    o_lab, o_cen = co.slic(img, step, nc)
    assert np.array_equal(r_lab, o_lab)
    assert r_cen.shape == o_cen.shape
    assert np.array_equal(r_cen.view(np.uint64), o_cen.view(np.uint64))
    assert r_cnt.shape[0] == r_cen.shape[0]


def test_slic_reference_flat_image_has_empty_clusters():
    """A constant image: ties everywhere (lowest centre index wins, strict <) and 0/0 centres for clusters that lose all pixels."""
    img = np.full((48, 64, 3), 128, np.uint8)
    r_lab, r_cen, _ = ro.generate_superpixels(img, 9, 40)
    o_lab, o_cen = co.slic(img, 9, 40)
    assert np.array_equal(r_lab, o_lab)
    assert np.array_equal(r_cen.view(np.uint64), o_cen.view(np.uint64))
