"""TEST INFRASTRUCTURE ONLY -- cv2 transliteration of the reference hot path.

This file is the *pinning* oracle: it restates the reference's C++ line by line and calls the
very same third-party kernels the reference calls (OpenCV ``imgproc``; the reference pins no
version, this oracle pins ``opencv-python-headless==4.13.0.92``).  It exists to
  (1) generate the golden vectors under ``tests/golden/`` (``oracle/make_golden.py``),
  (2) validate the plain-C restatement ``oracle/dcmt_oracle.c`` (which is what travels to the
      GPU box as the checker), and
  (3) serve as the "reference CPU path" timed by ``bench.py``'s ``cpu_baseline`` leg.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import it.  The product path (``depth_completion_mt_b200``) never does.

Parity status: the reference ships no golden vectors and no tests, and its own build needs
OpenCV / Eigen / PCL headers that the image lacks (``img_completion.h`` is missing from its
repo as well).  The pin is therefore the reference's OWN SOURCES compiled against a stand-in
header (``oracle/refshim``, ``oracle/ref_oracle.py`` -> ``oracle/_ref/libdcmt_ref.so``) with the
imgproc calls forwarded to this same OpenCV build: ``tests/test_reference_build.py`` requires
this transliteration, the C restatement and the golden vectors to equal its output bit for bit.

Reference files followed (relative to /root/reference):
  src/DC_lidar_only/img_completion.cpp:17-204          -> img_completion
  src/DC_lidar_camera/img_completion_lc.cpp:34-203     -> interpolate_with_superpixels
  src/DC_stereo_lidar/main_sl.cpp:715-885,1165-1253    -> stereo_* functions
"""
from __future__ import annotations

import numpy as np

try:  # cv2 is present in the build image; tests that need it skip when it is not.
    import cv2

    HAVE_CV2 = True
except Exception:  # pragma: no cover
    cv2 = None
    HAVE_CV2 = False

F32 = np.float32
MAX_DEPTH = F32(100.0)
# img_completion.cpp:59 compares a float against the double literal 0.1:
#   depth > 0.1  <=>  depth >= 0.1f  (0.1f is the smallest float above 0.1)
#   depth < 0.1  <=>  depth <  0.1f
THR = F32(0.1)


def diamond_kernel_as_read_by_opencv() -> np.ndarray:
    """img_completion.cpp:71-77: ``int d[5][5]`` wrapped as a CV_8UC1 Mat -> OpenCV reads the
    first 25 *bytes* of the int array.  On little-endian only (1,3) and (4,4) are non-zero."""
    d = np.array(
        [[0, 0, 1, 0, 0], [0, 1, 1, 1, 0], [1, 1, 1, 1, 1], [0, 1, 1, 1, 0], [0, 0, 1, 0, 0]],
        np.int32,
    )
    return d.view(np.uint8).ravel()[:25].reshape(5, 5).copy()


def _valid(d: np.ndarray) -> np.ndarray:
    return d >= THR


def _hole(d: np.ndarray) -> np.ndarray:
    return d < THR


def _invert(d: np.ndarray) -> None:
    # img_completion.cpp:55-67 / :191-202   d = 100 - d where d > 0.1 (float subtraction)
    m = _valid(d)
    d[m] = MAX_DEPTH - d[m]


def _column_extrapolation(d: np.ndarray) -> None:
    """img_completion.cpp:103-129 (always on: ``int densify = true``)."""
    rows, cols = d.shape
    v = _valid(d)
    any_v = v.any(axis=0)
    # max_index: last valid row (0 if none); min_index: first valid row (rows-1 if none)
    last = np.where(any_v, rows - 1 - np.argmax(v[::-1], axis=0), 0)
    first = np.where(any_v, np.argmax(v, axis=0), rows - 1)
    cidx = np.arange(cols)
    max_val = np.where(any_v, d[last, cidx], F32(-1.0)).astype(F32)
    min_val = np.where(any_v, d[first, cidx], F32(100.0)).astype(F32)
    r = np.arange(rows)[:, None]
    below = r >= last[None, :]
    above = r <= first[None, :]
    out = np.where(below, max_val[None, :], d)
    out = np.where(above, min_val[None, :], out)  # second loop wins where both apply
    d[...] = out


def _tail(d: np.ndarray, blur_type: str, stats: dict | None, gaussian_unconditional: bool) -> np.ndarray:
    """img_completion.cpp:88-202 == img_completion_lc.cpp:105-202 (the latter ignores blur_type)."""
    k7 = np.ones((7, 7), np.uint8)
    k31 = np.ones((31, 31), np.uint8)
    # :88-100  7x7 dilate, fill holes
    s = cv2.dilate(d, k7)
    h = _hole(d)
    d[h] = s[h]
    # :103-129
    _column_extrapolation(d)
    # :131-144  first large fill
    s = cv2.dilate(d, k31)
    h = _hole(d)
    d[h] = s[h]
    # :146-166  while loop, count taken before the fill, >= 1 pass
    passes = 0
    last_count = -1
    while True:
        s = cv2.dilate(d, k31)
        h = _hole(d)
        count = int(h.sum())
        d[h] = s[h]
        passes += 1
        last_count = count
        if count == 0:
            break
        if passes > 10000:  # the reference would spin forever; cannot happen after A5
            raise RuntimeError("31x31 fill does not terminate")
    if stats is not None:
        stats["loop_passes"] = passes
    # :170
    d = cv2.medianBlur(d, 5)
    if gaussian_unconditional or blur_type == "gaussian":
        # :176-189
        g = cv2.GaussianBlur(d, (5, 5), 0)
        m = _valid(d)
        d[m] = g[m]
    elif blur_type == "bilateral":
        # :172-175 -- the reference passes src==dst and OpenCV asserts; the evident intent is the
        # out-of-place call.
        d = cv2.bilateralFilter(d, 5, 1.5, 2.0)
    # :191-202
    _invert(d)
    return d


def img_completion(sparse: np.ndarray, blur_type: str = "gaussian", stats: dict | None = None) -> np.ndarray:
    """src/DC_lidar_only/img_completion.cpp:17-204.  ``extr`` is ignored by the reference."""
    assert sparse.dtype == np.float32 and sparse.ndim == 2
    d = sparse.copy()  # :27 clone
    _invert(d)  # :55-67
    d = cv2.dilate(d, diamond_kernel_as_read_by_opencv())  # :71-80
    d = cv2.morphologyEx(d, cv2.MORPH_CLOSE, np.ones((5, 5), np.uint8))  # :84-85
    return _tail(d, blur_type, stats, gaussian_unconditional=False)


def interpolate_with_superpixels(
    sparse: np.ndarray,
    labels: np.ndarray,
    n_clusters: int,
    blur_type: str = "gaussian",
    use_superpixel: int = 1,
    stats: dict | None = None,
) -> np.ndarray:
    """src/DC_lidar_camera/img_completion_lc.cpp:34-203.

    ``labels`` is row-major [row][col] int32 (the reference's ``slic.clusters`` is [col][row],
    img_completion_lc.cpp:83); ``n_clusters`` is ``slic.centers.size()``.
    """
    assert sparse.dtype == np.float32 and labels.shape == sparse.shape
    d = sparse.copy()
    _invert(d)  # :45-52
    dk = diamond_kernel_as_read_by_opencv()
    k5 = np.ones((5, 5), np.uint8)
    if use_superpixel == 0:
        d = cv2.dilate(d, dk)  # :59-64
        d = cv2.morphologyEx(d, cv2.MORPH_CLOSE, k5)
    else:
        for c in range(int(n_clusters)):  # :78-103
            mask = labels == c
            if not mask.any():
                continue  # region.copyTo(dense, empty mask) is a no-op
            region = np.zeros_like(d)  # copyTo into a fresh Mat zero-fills outside the mask
            region[mask] = d[mask]
            region = cv2.dilate(region, dk)
            region = cv2.morphologyEx(region, cv2.MORPH_CLOSE, k5)
            d[mask] = region[mask]
    return _tail(d, blur_type, stats, gaussian_unconditional=True)


# --------------------------------------------------------------------------------------
# Stereo refinement (main_sl.cpp).  Plain numpy float32, source operation order, no cv2
# except the final GaussianBlur.  Reads at column index == cols / row index == rows are
# undefined behaviour in the reference (:763-779, over-allocated Mats :1165,1169); they are
# DEFINED as 0 here.
# --------------------------------------------------------------------------------------
BASELINE = F32(0.54)
FOCAL = F32(9.597910e02)


def measurement_derivatives(val: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """main_sl.cpp:715-745  central differences on the interior, 0 on the 1-px border.
    ``.5 * v`` is double arithmetic rounded to float on store; both products are exact."""
    v = val.astype(np.float64)
    dx = np.zeros_like(v)
    dy = np.zeros_like(v)
    dx[1:-1, 1:-1] = 0.5 * v[1:-1, 2:] - 0.5 * v[1:-1, :-2]
    dy[1:-1, 1:-1] = 0.5 * v[2:, 1:-1] - 0.5 * v[:-2, 1:-1]
    return dx.astype(F32), dy.astype(F32)


def get_initial_disparity(depth: np.ndarray) -> np.ndarray:
    """main_sl.cpp:846-861 (disparity map pre-zeroed at :1191)."""
    bf = F32(BASELINE * FOCAL)
    out = np.zeros_like(depth, dtype=F32)
    m = depth > 0
    out[m] = bf / depth[m]
    return out


def optimize_IG(
    val_l: np.ndarray,
    val_r: np.ndarray,
    disp: np.ndarray,
    num_iterations: int = 4,
    damp_factor: float = 500.0,
    err_clip: float = 255.0,
) -> np.ndarray:
    """main_sl.cpp:804-843 with calculateObservationDerivatives :747-801 inlined (vectorised;
    pixels are independent).  OFFICIAL variant: damp 1370, clip 221 (main_sl_OFFICIAL.cpp:832-876)."""
    rows, cols = val_l.shape
    dxr, _ = measurement_derivatives(val_r)
    # zero-extended planes so that index == cols / rows reads the defined 0
    vr = np.zeros((rows + 2, cols + 2), F32)
    gr = np.zeros((rows + 2, cols + 2), F32)
    vr[:rows, :cols] = val_r
    gr[:rows, :cols] = dxr
    disp = disp.astype(F32).copy()
    jj = np.arange(cols, dtype=F32)[None, :].repeat(rows, 0)
    ii = np.arange(rows)[:, None].repeat(cols, 1)
    damp = F32(damp_factor)
    clip = F32(err_clip)
    for _ in range(num_iterations):
        c = (jj - disp).astype(F32)  # float pixel_right = j - disp
        c0 = np.trunc(c.astype(np.float64) + 0.5).astype(np.int64)  # (int)(c + 0.5), double add
        r0 = ii  # r is integral, (int)(r + 0.5) == r
        ok = ~((c0 < 0) | (c0 > cols) | (c0 + 1 > cols))  # row tests never fire for i < rows
        ok &= disp != 0
        c0c = np.clip(c0, 0, cols)
        dc = (c - c0c.astype(F32)).astype(F32)
        dc1 = (1.0 - dc.astype(np.float64)).astype(F32)  # ``1. - dc`` in double, stored as float
        p00 = vr[r0, c0c]
        p01 = vr[r0, c0c + 1]
        value = (p00 * dc1).astype(F32) + (p01 * dc).astype(F32)  # dr == 0, dr1 == 1
        g00 = gr[r0, c0c]
        g01 = gr[r0, c0c + 1]
        gx = (g00 * dc1).astype(F32) + (g01 * dc).astype(F32)
        err = (value - val_l).astype(F32)
        err = np.minimum(np.maximum(err, -clip), clip)
        jcr = (F32(-1.0) * gx).astype(F32)
        hh = (jcr * jcr).astype(F32) + damp
        b = (jcr * err).astype(F32)
        dd = (-b / hh).astype(F32)
        disp = np.where(ok, (disp + dd).astype(F32), disp)
    return disp


def retrieve_optimized_depth(disp: np.ndarray, depth_clip: float = 100.0) -> np.ndarray:
    """main_sl.cpp:863-885 (output pre-zeroed at :1244)."""
    bf = F32(BASELINE * FOCAL)
    out = np.zeros_like(disp, dtype=F32)
    m = disp > 0
    dep = bf / disp[m]
    out[m] = np.minimum(dep, F32(depth_clip))  # if (depth > 100) depth = 100
    return out


def stereo_refine(
    depth_ig: np.ndarray,
    left_gray: np.ndarray,
    right_gray: np.ndarray,
    num_iterations: int = 4,
    damp_factor: float = 500.0,
    err_clip: float = 255.0,
    depth_clip: float = 100.0,
    final_gauss: bool = True,
) -> np.ndarray:
    """main_sl.cpp:1165-1253: gray u8 -> entries, initial disparity, GN refinement, depth, blur."""
    vl = left_gray.astype(F32)
    vr = right_gray.astype(F32)
    disp = get_initial_disparity(depth_ig)
    disp = optimize_IG(vl, vr, disp, num_iterations, damp_factor, err_clip)
    dep = retrieve_optimized_depth(disp, depth_clip)
    if final_gauss:
        dep = cv2.GaussianBlur(dep, (5, 5), 0)  # :1253
    return dep
