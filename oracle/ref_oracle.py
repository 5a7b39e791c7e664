"""TEST INFRASTRUCTURE ONLY -- the reference's own code as the checker of the checker.

oracle/_ref/libdcmt_ref.so holds img_completion (src/DC_lidar_only/img_completion.cpp:17-204),
interpolate_with_superpixels (src/DC_lidar_camera/img_completion_lc.cpp:34-203) and Slic
(src/DC_lidar_camera/slic.cpp) compiled UNMODIFIED from /root/reference against the stand-in OpenCV header in
oracle/refshim/ (recipe: oracle/Makefile).  The stand-in implements containers only; the five imgproc calls the
reference makes are forwarded, through the callbacks registered here, to cv2 -- the same OpenCV 4.13 the
transliteration (oracle/cv2_oracle.py) uses.  So: control flow, scalar loops and buffer handling are the reference's
compiled code, the filters are the real library.

It pins the restatements (cv2_oracle.py, dcmt_oracle.c) to the reference itself (tests/test_reference_build.py) and can
serve as bench.py's CPU arm (`cpu_baseline.kind = "reference"`).  /root/reference exists only in the build container:
there the library is (re)built on demand; elsewhere the prebuilt file that travelled with the snapshot is used.
Never imported by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libdcmt_ref.so")
_REF = os.environ.get("DCMT_REFERENCE_DIR", "/root/reference")
_SHIM = [os.path.join(_HERE, "refshim", n) for n in ("refshim.cpp", "refshim_ranges.cpp", "refshim_front.cpp", "extract_ranges.sh",
                                                     "img_completion.h", os.path.join("opencv2", "opencv.hpp"), os.path.join("Eigen", "Dense"))]
BLUR = {"none": 0, "gaussian": 1, "bilateral": 2}

_MORPH_CB = C.CFUNCTYPE(C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int)
_BLUR_CB = C.CFUNCTYPE(C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double)
_CV_TYPES = {0: np.uint8, 5: np.float32}  # CV_8UC1, CV_32FC1


class ReferenceUnavailable(RuntimeError):
    pass


def have_sources() -> bool:
    return os.path.isfile(os.path.join(_REF, "src", "DC_lidar_only", "img_completion.cpp"))


def build(force: bool = False) -> str:
    """(Re)build oracle/_ref/libdcmt_ref.so where the reference sources exist; otherwise return the prebuilt file."""
    if have_sources():
        stale = force or not os.path.exists(_SO) or any(os.path.getmtime(_SO) < os.path.getmtime(p) for p in _SHIM)
        if stale:
            subprocess.run(["make", "-C", _HERE, "-B", "_ref/libdcmt_ref.so", f"REF={_REF}"], check=True, capture_output=True)
    if not os.path.exists(_SO):
        raise ReferenceUnavailable(f"{_SO} is missing and {_REF} is not present to build it from")
    return _SO


def available() -> bool:
    try:
        build()
        import cv2  # noqa: F401
        return True
    except Exception:
        return False


def _view(ptr, rows, cols, cv_type):
    dt = _CV_TYPES[cv_type]
    n = rows * cols
    buf = (C.c_ubyte * (n * np.dtype(dt).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dt, count=n).reshape(rows, cols)


def _morph(op, src, dst, rows, cols, cv_type, kernel, krows, kcols):
    try:
        import cv2
        s = _view(src, rows, cols, cv_type)
        k = _view(kernel, krows, kcols, 0)
        if op == 1:
            r = cv2.dilate(s, k)
        elif op == 3:
            r = cv2.morphologyEx(s, cv2.MORPH_CLOSE, k)
        else:
            return 2
        _view(dst, rows, cols, cv_type)[...] = r
        return 0
    except Exception:  # an exception must not cross the C boundary
        return 1


def _blur(kind, src, dst, rows, cols, cv_type, ksize, p0, p1):
    try:
        import cv2
        s = _view(src, rows, cols, cv_type)
        if kind == 0:
            r = cv2.medianBlur(s, ksize)
        elif kind == 1:
            r = cv2.GaussianBlur(s, (ksize, ksize), p0)
        elif kind == 2:
            r = cv2.bilateralFilter(s, ksize, p0, p1)
        elif kind == 3:  # cv::normalize(src, dst, alpha, beta, NORM_MINMAX) (main_sl.cpp:523)
            r = cv2.normalize(s, None, p0, p1, cv2.NORM_MINMAX)
        else:
            return 2
        _view(dst, rows, cols, cv_type)[...] = r
        return 0
    except Exception:
        return 1


_lib = None
_keep = []  # the ctypes callback objects must outlive the library's use of them


def lib():
    global _lib
    if _lib is None:
        import cv2
        cv2.setNumThreads(1)
        L = C.CDLL(build())
        m, b = _MORPH_CB(_morph), _BLUR_CB(_blur)
        _keep.extend([m, b])
        L.dcmt_ref_set_callbacks(m, b)
        _lib = L
    return _lib


def _err():
    return C.create_string_buffer(256)


def img_completion(sparse, blur_type="gaussian"):
    """The reference's img_completion on one float32 frame.  Raises RuntimeError with the reference's exception text
    (blur_type="bilateral": OpenCV's in-place assertion, img_completion.cpp:174)."""
    s = np.ascontiguousarray(sparse, dtype=np.float32)
    rows, cols = s.shape
    out = np.empty_like(s)
    e = _err()
    rc = lib().dcmt_ref_img_completion(s.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), rows, cols, BLUR[blur_type], e, 256)
    if rc:
        raise RuntimeError(e.value.decode(errors="replace"))
    return out


def interpolate_with_superpixels(labels, sparse, blur_type="gaussian", use_superpixel=1, n_clusters=None):
    """The reference's interpolate_with_superpixels; labels row-major [row][col] (the shim fills Slic::clusters[col][row]),
    n_clusters = slic.centers.size() (default: max label + 1)."""
    s = np.ascontiguousarray(sparse, dtype=np.float32)
    lab = np.ascontiguousarray(labels, dtype=np.int32)
    rows, cols = s.shape
    if n_clusters is None:
        n_clusters = int(lab.max()) + 1 if lab.size else 0
    out = np.empty_like(s)
    e = _err()
    rc = lib().dcmt_ref_interpolate_with_superpixels(s.ctypes.data_as(C.c_void_p), lab.ctypes.data_as(C.c_void_p), int(n_clusters),
                                                     out.ctypes.data_as(C.c_void_p), rows, cols, BLUR[blur_type], int(use_superpixel), e, 256)
    if rc:
        raise RuntimeError(e.value.decode(errors="replace"))
    return out


def generate_superpixels(lab_image, step, nc):
    """Slic::generate_superpixels on an H x W x 3 uint8 image -> (labels int32 H x W, centers float64 K x 5, counts int32 K)."""
    img = np.ascontiguousarray(lab_image, dtype=np.uint8)
    rows, cols, ch = img.shape
    assert ch == 3
    cap = (rows // max(int(step), 1) + 2) * (cols // max(int(step), 1) + 2) + 8
    labels = np.empty((rows, cols), np.int32)
    centers = np.zeros((cap, 5), np.float64)
    counts = np.zeros(cap, np.int32)
    e = _err()
    n = lib().dcmt_ref_slic(img.ctypes.data_as(C.c_void_p), rows, cols, int(step), int(nc), labels.ctypes.data_as(C.c_void_p),
                            centers.ctypes.data_as(C.c_void_p), counts.ctypes.data_as(C.c_void_p), cap, e, 256)
    if n < 0:
        raise RuntimeError(e.value.decode(errors="replace"))
    assert n <= cap
    return labels, centers[:n].copy(), counts[:n].copy()


def call_counts():
    """How many times the reference called (dilate, morphologyEx, medianBlur, GaussianBlur, bilateralFilter) so far."""
    a = (C.c_longlong * 5)()
    lib().dcmt_ref_call_counts(a)
    return tuple(a)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def measurement_derivatives(value):
    """calculateMeasuementDerivatives (main_sl.cpp:715-745) -> (derivative.x(), derivative.y()) planes."""
    v = np.ascontiguousarray(value, dtype=np.float32)
    dx, dy = np.empty_like(v), np.empty_like(v)
    lib().dcmt_ref_measurement_derivatives(_p(v), _p(dx), _p(dy), *v.shape)
    return dx, dy


def get_initial_disparity(depth):
    d = np.ascontiguousarray(depth, dtype=np.float32)
    out = np.empty_like(d)
    lib().dcmt_ref_get_initial_disparity(_p(d), _p(out), *d.shape)
    return out


def retrieve_optimized_depth(disp):
    d = np.ascontiguousarray(disp, dtype=np.float32)
    out = np.empty_like(d)
    lib().dcmt_ref_retrieve_optimized_depth(_p(d), _p(out), *d.shape)
    return out


def optimize_IG(value_left, value_right, disp):
    """optimize_IG (main_sl.cpp:804-843) on float value planes; returns the refined disparity (input untouched)."""
    l = np.ascontiguousarray(value_left, dtype=np.float32)
    r = np.ascontiguousarray(value_right, dtype=np.float32)
    d = np.array(disp, dtype=np.float32, order="C", copy=True)
    lib().dcmt_ref_optimize_IG(_p(l), _p(r), _p(d), *d.shape)
    return d


def stereo_refine(depth_ig, left_gray, right_gray, final_gauss=True, return_disp=False):
    """The chain of main_sl.cpp:1165-1253 with the constants hard-coded there (4 iterations, damping 500, clips 255 / 100)."""
    d = np.ascontiguousarray(depth_ig, dtype=np.float32)
    lg = np.ascontiguousarray(left_gray, dtype=np.uint8)
    rg = np.ascontiguousarray(right_gray, dtype=np.uint8)
    out, disp = np.empty_like(d), np.empty_like(d)
    e = _err()
    rc = lib().dcmt_ref_stereo_chain(_p(d), _p(lg), _p(rg), _p(out), _p(disp), d.shape[0], d.shape[1], int(bool(final_gauss)), e, 256)
    if rc:
        raise RuntimeError(e.value.decode(errors="replace"))
    return (out, disp) if return_disp else out


def evaluate(gt, r, variant):
    """variant "lidar_only" (main.cpp:16-34) -> mean signed error; "lidar_camera" (main_lc.cpp:85-116) -> (rmse, mae);
    "stereo_lidar" (main_sl.cpp:1031-1061) -> (mae, rmse).  float32, exactly as the reference accumulates them."""
    g = np.ascontiguousarray(gt, dtype=np.float32)
    v = np.ascontiguousarray(r, dtype=np.float32)
    code = {"lidar_only": 0, "lidar_camera": 1, "stereo_lidar": 2}[variant]
    out = np.zeros(2, np.float32)
    rc = lib().dcmt_ref_evaluate(code, _p(g), _p(v), g.shape[0], g.shape[1], _p(out))
    assert rc == 0
    return np.float32(out[0]) if code == 0 else (np.float32(out[0]), np.float32(out[1]))


def lidar_project(points, T, P, rows, cols):
    """The projection loops + cv::normalize(0, 80) of withSuperPixels (main_sl.cpp:474-523), compiled from the reference's
    own lines (refshim_front.cpp).  points: n x 4 float32; T 4x4, P 3x4.  Returns (projected, normalized, n_projected)."""
    pts = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 4)
    t = np.ascontiguousarray(T, dtype=np.float32).reshape(16)
    p = np.ascontiguousarray(P, dtype=np.float32).reshape(12)
    proj = np.empty((rows, cols), np.float32)
    nrm = np.empty((rows, cols), np.float32)
    n = C.c_int(0)
    e = _err()
    rc = lib().dcmt_ref_lidar_project(_p(pts), len(pts), _p(t), _p(p), rows, cols, _p(proj), _p(nrm), C.byref(n), e, 256)
    if rc:
        raise RuntimeError(e.value.decode(errors="replace"))
    return proj, nrm, n.value


def write_M(path, mat):
    """write_M (utils.cpp:39-58) of a CV_32FC1 matrix."""
    m = np.ascontiguousarray(mat, dtype=np.float32)
    lib().dcmt_ref_write_M(str(path).encode(), _p(m), *m.shape)


def read_M(path):
    """read_M (utils.cpp:15-37): the CV_32FC1 matrix of a .bin file, or None."""
    rows, cols, typ = C.c_int(0), C.c_int(0), C.c_int(0)
    cap = os.path.getsize(path) // 4
    out = np.empty(max(cap, 1), np.float32)
    n = lib().dcmt_ref_read_M(str(path).encode(), C.byref(rows), C.byref(cols), C.byref(typ), _p(out), int(cap))
    if n == 0 or typ.value != 5:
        return None
    return out[:n].reshape(rows.value, cols.value).copy()
