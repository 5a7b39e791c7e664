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


# ---------------------------------------------------------------------------------------------------------------------
# stereo refinement (main_sl.cpp:23-26, 715-885, 1165-1253) and the evaluation loops, from the reference's own lines


@pytest.mark.parametrize("shape", [(352, 1216), (64, 100), (9, 12), (3, 3)])
def test_stereo_chain_reference_vs_restatements(shape):
    """a3-a9: EntryType matrices, derivatives, initial disparity, 4 GN iterations, depth retrieval: bit for bit; with the
    final cv::GaussianBlur (float data) within 1e-4 for the C port and bit-equal for the transliteration (same OpenCV)."""
    dig, left, right = synth.stereo_pair(5, shape[0], shape[1])
    ref, ref_disp = ro.stereo_refine(dig, left, right, final_gauss=False, return_disp=True)
    out, disp = co.stereo_refine(dig, left, right, final_gauss=False, return_disp=True)
    assert_bit_equal(disp, ref_disp, f"{shape} disparity after 4 iterations")
    assert_bit_equal(out, ref, f"{shape} depth")
    blurred = ro.stereo_refine(dig, left, right, final_gauss=True)
    assert np.abs(co.stereo_refine(dig, left, right) - blurred).max() <= 1e-4
    if min(shape) >= 5:
        assert_bit_equal(np.asarray(cvo.stereo_refine(dig, left, right)), blurred, f"{shape} transliteration")


def test_stereo_functions_one_by_one():
    dig, left, right = synth.stereo_pair(6, 48, 80)
    lf, rf = left.astype(np.float32), right.astype(np.float32)
    for a, b, what in zip(co.measurement_derivatives(rf), ro.measurement_derivatives(rf), ("dx", "dy")):
        assert_bit_equal(a, b, what)
    d0 = ro.get_initial_disparity(dig)
    assert_bit_equal(np.asarray(cvo.get_initial_disparity(dig)), d0, "get_initial_disparity")
    d4 = ro.optimize_IG(lf, rf, d0)
    _, disp = co.stereo_refine(dig, left, right, final_gauss=False, return_disp=True)
    assert_bit_equal(disp, d4, "optimize_IG")
    assert_bit_equal(np.asarray(cvo.retrieve_optimized_depth(d4)), ro.retrieve_optimized_depth(d4), "retrieve_optimized_depth")


def test_stereo_zero_and_far_depths():
    """depth 0 -> disparity 0 -> never refined, output 0; tiny disparities clip at 100 m (main_sl.cpp:875-878)."""
    dig, left, right = synth.stereo_pair(8, 40, 64)
    dig = dig.copy()
    dig[:5] = 0.0
    dig[5:8] = 5000.0
    ref = ro.stereo_refine(dig, left, right, final_gauss=False)
    assert_bit_equal(co.stereo_refine(dig, left, right, final_gauss=False), ref, "zero / far depths")
    assert (ref[:5] == 0).all() and (ref[5:8] <= 100.0).all()


def test_reference_reproduces_stereo_golden(golden):
    g = golden["stereo"]
    for name in sorted({k.split("__")[0] for k in g.files}):
        dig, left, right = g[name + "__depth_ig"], g[name + "__left"], g[name + "__right"]
        out, disp = ro.stereo_refine(dig, left, right, final_gauss=False, return_disp=True)
        assert_bit_equal(out, g[name + "__default_nogauss"], f"{name} depth")
        assert_bit_equal(disp, g[name + "__disp4"], f"{name} disparity")
        assert_bit_equal(ro.stereo_refine(dig, left, right), g[name + "__default"], f"{name} with the final Gaussian")


@pytest.mark.parametrize("shape", [(37, 53), (352, 1216)])
def test_evaluation_loops_reference_vs_restatement(shape):
    """f3: the three float32 raster-order loops (int tolerance 0 / (int)0.1 = 0 / 2)."""
    rng = np.random.default_rng(11)
    gt = np.where(rng.random(shape) < 0.3, rng.uniform(0.5, 80, shape), 0).astype(np.float32)
    dense = (rng.uniform(0.0, 85, shape) * (rng.random(shape) < 0.9)).astype(np.float32)
    if shape[0] < 100:  # main.cpp:29 prints every pixel pair; keep that one small
        assert np.float32(co.evaluate(gt, dense, 0, 0)["mean_err"]) == ro.evaluate(gt, dense, "lidar_only")
    rmse, mae = ro.evaluate(gt, dense, "lidar_camera")
    o = co.evaluate(gt, dense, 0, 1)
    assert (np.float32(o["rmse"]), np.float32(o["mae"])) == (rmse, mae)
    mae2, rmse2 = ro.evaluate(gt, dense, "stereo_lidar")
    o = co.evaluate(gt, dense, 2, 1)
    assert (np.float32(o["mae"]), np.float32(o["rmse"])) == (mae2, rmse2)


# ---------------------------------------------------------------------------------------------------------------------
# the product against the reference build directly (B200, through the C ABI)


@pytest.mark.gpu
def test_gpu_img_completion_equals_reference_build(gpu_lib):
    """Fused strict-q8 kernels (device and host entry points, float32 and KITTI uint16 input) and the generic float32
    pipeline == the reference's compiled img_completion at 352 x 1216, every bit."""
    import torch

    from depth_completion_mt_b200 import api

    for frame, density, kitti_like in ((0, 0.05, False), (1, 0.05, True), (2, 0.01, False), (3, 0.2, False)):
        d16 = synth.sparse_depth_q8(frame, density=density, kitti_like=kitti_like)
        s = d16.astype(np.float32) / np.float32(256)
        for bt in ("gaussian", "none"):
            ref = ro.img_completion(s, bt)
            assert_bit_equal(api.img_completion(torch.from_numpy(s).cuda(), False, bt, lib=gpu_lib).cpu().numpy(), ref, f"fused {frame} {bt}")
            assert_bit_equal(api.img_completion(torch.from_numpy(s).cuda(), False, bt, path="generic", lib=gpu_lib).cpu().numpy(), ref, f"generic {frame} {bt}")
        assert_bit_equal(api.img_completion(s, False, "gaussian", lib=gpu_lib), ro.img_completion(s, "gaussian"), f"host entry point {frame}")
        assert_bit_equal(api.img_completion(torch.from_numpy(d16).cuda(), False, "gaussian", lib=gpu_lib).cpu().numpy(),
                         ro.img_completion(s, "gaussian"), f"uint16 input {frame}")


@pytest.mark.gpu
def test_gpu_guided_and_stereo_equal_reference_build(gpu_lib):
    import torch

    from depth_completion_mt_b200 import api

    rows, cols = 64, 96  # the reference runs three full-frame morphology calls per superpixel: keep it small
    s = synth.sparse_depth(9, rows, cols, 0.08)
    lab, k = synth.superpixel_labels(9, rows, cols, step=12)
    lab = lab.copy()
    lab[3:6, 40:50] = -1
    for sp in (1, 0):
        got = api.interpolate_with_superpixels(torch.from_numpy(lab).cuda(), torch.from_numpy(s).cuda(), "gaussian", sp, n_clusters=k, lib=gpu_lib)
        assert_bit_equal(got.cpu().numpy(), ro.interpolate_with_superpixels(lab, s, use_superpixel=sp, n_clusters=k), f"guided sp={sp}")
    dig, left, right = synth.stereo_pair(4)
    prm = api.stereo_params(final_gauss=0, lib=gpu_lib)
    got = api.stereo_refine(torch.from_numpy(dig).cuda(), torch.from_numpy(left).cuda(), torch.from_numpy(right).cuda(), prm, lib=gpu_lib)
    assert_bit_equal(got.cpu().numpy(), ro.stereo_refine(dig, left, right, final_gauss=False), "stereo refinement at 352 x 1216")
    got = api.stereo_refine(torch.from_numpy(dig).cuda(), torch.from_numpy(left).cuda(), torch.from_numpy(right).cuda(), lib=gpu_lib)
    assert np.abs(got.cpu().numpy() - ro.stereo_refine(dig, left, right)).max() <= 1e-4  # final float Gaussian (SURVEY 8c tolerance)


@pytest.mark.gpu
def test_gpu_slic_equals_reference_build(gpu_lib):
    import torch

    from depth_completion_mt_b200 import api

    img = synth.lab_image(1)
    labels, centers = api.generate_superpixels(torch.from_numpy(img).cuda(), 18, 40, return_centers=True, lib=gpu_lib)
    r_lab, r_cen, _ = ro.generate_superpixels(img, 18, 40)
    assert np.array_equal(labels.cpu().numpy(), r_lab)
    assert np.array_equal(np.ascontiguousarray(centers.cpu().numpy(), dtype=np.float64).view(np.uint64), r_cen.view(np.uint64))


# ---- the callers either side of the path, pinned to the reference's own lines ----------------------
T_KITTI, P_KITTI = synth.KITTI_T_VELO_TO_CAM, synth.KITTI_P_RECT_02


@pytest.mark.parametrize("rows,cols,n", [(352, 1216, 120000), (375, 1242, 60000), (64, 200, 20000), (48, 160, 30000)])
def test_projection_loop_reference_vs_restatement(rows, cols, n):
    """main_sl.cpp:474-523 compiled from the reference (two PCL types and the Eigen product are stand-ins, see
    refshim_front.cpp / refshim/Eigen/Dense) == the C restatement the GPU kernel is tested against, bit for bit:
    projected depth image, its cv::normalize(0, 80) and the count of projected points."""
    pts = synth.velodyne_cloud({1216: 0, 1242: 5, 200: 0, 160: 1}[cols], n)
    if cols == 160:  # many points per pixel: last writer in file order wins
        pts = np.concatenate([pts, pts[::-1] * np.float32(1.0001), pts[: n // 2]])
    rp, rn, rc = ro.lidar_project(pts, T_KITTI, P_KITTI, rows, cols)
    cp, cn, cc = co.lidar_project(pts, T_KITTI, P_KITTI, rows, cols)
    assert rc == cc and (rc > 1000 or cols < 1000)
    assert_bit_equal(cp, rp, "projected_depths")
    assert_bit_equal(cn, rn, "normalized_depths")


def test_projection_loop_degenerate_clouds():
    behind = synth.velodyne_cloud(3, 2000)
    behind[:, 0] = -np.abs(behind[:, 0]) - 1.0
    for pts in (np.zeros((0, 4), np.float32), behind):
        rp, rn, rc = ro.lidar_project(pts, T_KITTI, P_KITTI, 40, 60)
        cp, cn, cc = co.lidar_project(pts, T_KITTI, P_KITTI, 40, 60)
        assert rc == cc == 0
        assert_bit_equal(cp, rp, "projected_depths (empty)")
        assert_bit_equal(cn, rn, "normalized_depths (empty)")


def test_raw_mat_format_reference_vs_package(tmp_path):
    """utils.cpp:15-58 (read_M / write_M, compiled from the reference) and depth_completion_mt_b200.api.read_M / write_M
    produce and accept the same bytes."""
    from depth_completion_mt_b200 import api

    m = synth.sparse_depth(4, 37, 53, 0.3)
    ro.write_M(tmp_path / "ref.bin", m)
    api.write_M(tmp_path / "pkg.bin", m)
    assert open(tmp_path / "ref.bin", "rb").read() == open(tmp_path / "pkg.bin", "rb").read()
    assert_bit_equal(api.read_M(tmp_path / "ref.bin"), m, "package reads the reference's file")
    assert_bit_equal(ro.read_M(tmp_path / "pkg.bin"), m, "the reference reads the package's file")
