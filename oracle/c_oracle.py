"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/libdcmt_oracle.so (oracle/dcmt_oracle.c).

Used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as the CPU checker.
Never imported by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libdcmt_oracle.so")
BLUR = {"none": 0, "gaussian": 1, "bilateral": 2}
STAGE_NAMES = (
    "invert", "two_tap", "close5", "dilate7_fill", "column_extrapolation",
    "fill31_first", "fill31_loop", "median5", "blur", "final_invert",
)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "dcmt_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "libdcmt_oracle.so"], check=True, capture_output=True)
    return _SO


class StereoParams(C.Structure):
    _fields_ = [
        ("baseline", C.c_float), ("focal", C.c_float), ("damp_factor", C.c_float),
        ("err_clip", C.c_float), ("depth_clip", C.c_float),
        ("num_iterations", C.c_int32), ("final_gauss", C.c_int32),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.dcmt_oracle_n_stages.restype = C.c_int
    return _lib


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


def img_completion(sparse, blur_type="gaussian", stats=None, stages=False):
    s, sp = _f32(sparse)
    rows, cols = s.shape
    out = np.empty_like(s)
    st = np.zeros(4, np.int32)
    stg = np.empty((lib().dcmt_oracle_n_stages(), rows, cols), np.float32) if stages else None
    rc = lib().dcmt_oracle_img_completion_stages(
        sp, out.ctypes.data_as(C.POINTER(C.c_float)), rows, cols, BLUR[blur_type],
        st.ctypes.data_as(C.POINTER(C.c_int32)),
        stg.ctypes.data_as(C.POINTER(C.c_float)) if stages else None,
    )
    if rc:
        raise RuntimeError(f"oracle img_completion rc={rc}")
    if stats is not None:
        stats.update(loop_passes=int(st[0]), holes_before_loop=int(st[1]), holes_after_extrapolation=int(st[2]))
    return (out, stg) if stages else out


def interpolate_with_superpixels(sparse, labels, n_clusters, use_superpixel=1, literal=False, stats=None, stages=False):
    s, sp = _f32(sparse)
    rows, cols = s.shape
    lab = np.ascontiguousarray(labels, dtype=np.int32)
    out = np.empty_like(s)
    st = np.zeros(4, np.int32)
    stg = np.empty((lib().dcmt_oracle_n_stages(), rows, cols), np.float32) if stages else None
    rc = lib().dcmt_oracle_interpolate_with_superpixels(
        sp, lab.ctypes.data_as(C.POINTER(C.c_int32)), int(n_clusters),
        out.ctypes.data_as(C.POINTER(C.c_float)), rows, cols, 1, int(use_superpixel), int(bool(literal)),
        st.ctypes.data_as(C.POINTER(C.c_int32)),
        stg.ctypes.data_as(C.POINTER(C.c_float)) if stages else None,
    )
    if rc:
        raise RuntimeError(f"oracle interpolate_with_superpixels rc={rc}")
    if stats is not None:
        stats.update(loop_passes=int(st[0]), holes_before_loop=int(st[1]), holes_after_extrapolation=int(st[2]))
    return (out, stg) if stages else out


def stereo_refine(depth_ig, left_gray, right_gray, num_iterations=4, damp_factor=500.0, err_clip=255.0,
                  depth_clip=100.0, final_gauss=True, baseline=0.54, focal=9.597910e02, return_disp=False):
    d, dp = _f32(depth_ig)
    rows, cols = d.shape
    lg = np.ascontiguousarray(left_gray, dtype=np.uint8)
    rg = np.ascontiguousarray(right_gray, dtype=np.uint8)
    out = np.empty_like(d)
    disp = np.empty_like(d)
    prm = StereoParams(baseline, focal, damp_factor, err_clip, depth_clip, num_iterations, int(final_gauss))
    rc = lib().dcmt_oracle_stereo_refine(
        dp, lg.ctypes.data_as(C.POINTER(C.c_uint8)), rg.ctypes.data_as(C.POINTER(C.c_uint8)),
        out.ctypes.data_as(C.POINTER(C.c_float)), disp.ctypes.data_as(C.POINTER(C.c_float)),
        rows, cols, C.byref(prm),
    )
    if rc:
        raise RuntimeError(f"oracle stereo_refine rc={rc}")
    return (out, disp) if return_disp else out


def measurement_derivatives(val):
    v, vp = _f32(val)
    dx = np.empty_like(v)
    dy = np.empty_like(v)
    lib().dcmt_oracle_measurement_derivatives(
        vp, dx.ctypes.data_as(C.POINTER(C.c_float)), dy.ctypes.data_as(C.POINTER(C.c_float)), *v.shape)
    return dx, dy


def _op(name, src, *extra):
    s, sp = _f32(src)
    out = np.empty_like(s)
    getattr(lib(), name)(sp, out.ctypes.data_as(C.POINTER(C.c_float)), s.shape[0], s.shape[1], *extra)
    return out


def op_two_tap(src):
    return _op("dcmt_oracle_op_two_tap", src)


def op_box(src, k, is_max):
    return _op("dcmt_oracle_op_box", src, int(k), int(is_max))


def op_median5(src):
    return _op("dcmt_oracle_op_median5", src)


def op_gaussian5(src):
    return _op("dcmt_oracle_op_gaussian5", src)


def op_bilateral5(src):
    return _op("dcmt_oracle_op_bilateral5", src)


def op_column_extrapolation(src):
    s = np.array(src, dtype=np.float32, order="C", copy=True)
    lib().dcmt_oracle_op_column_extrapolation(s.ctypes.data_as(C.POINTER(C.c_float)), s.shape[0], s.shape[1])
    return s


def evaluate(gt, r, tolerance: int = 0, mode: int = 1):
    """Literal float32 raster-order evaluation loop (main.cpp:16-34 mode 0; main_lc.cpp:85-116 / main_sl.cpp:1031-1061
    mode 1).  Returns dict(mean_err, mae, rmse, count)."""
    g, gp = _f32(gt)
    v, vp = _f32(r)
    out = np.zeros(4, np.float32)
    lib().dcmt_oracle_evaluate(gp, vp, g.shape[0], g.shape[1], int(tolerance), int(mode), out.ctypes.data_as(C.POINTER(C.c_float)))
    return {"mean_err": float(out[0]), "mae": float(out[1]), "rmse": float(out[2]), "count": int(out[3])}


def lidar_project(points, T, P, rows, cols, norm=(0.0, 80.0)):
    """Literal main_sl.cpp:478-523 + cv::normalize: returns (projected, normalized, count)."""
    pts, pp = _f32(np.asarray(points, np.float32).reshape(-1, 4))
    Tm, tp = _f32(np.asarray(T, np.float32).reshape(4, 4))
    Pm, ppm = _f32(np.asarray(P, np.float32).reshape(3, 4))
    proj = np.empty((rows, cols), np.float32)
    nrm = np.empty((rows, cols), np.float32)
    fn = lib().dcmt_oracle_lidar_project
    fn.restype = C.c_int
    fn.argtypes = [C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.c_int, C.POINTER(C.c_float),
                   C.POINTER(C.c_float), C.c_float, C.c_float]
    cnt = fn(pp, pts.shape[0], tp, ppm, rows, cols, proj.ctypes.data_as(C.POINTER(C.c_float)), nrm.ctypes.data_as(C.POINTER(C.c_float)),
             float(norm[0]), float(norm[1]))
    return proj, nrm, int(cnt)


def slic(lab, step, nc, iterations=10):
    """Literal Slic::generate_superpixels (slic.cpp:101-182): returns (labels int32 [row][col], centers (K, 5) float64)."""
    lab = np.ascontiguousarray(lab, np.uint8)
    rows, cols = lab.shape[:2]
    labels = np.empty((rows, cols), np.int32)
    centers = np.empty((max(1, (rows // max(int(step), 1) + 2) * (cols // max(int(step), 1) + 2)), 5), np.float64)
    fn = lib().dcmt_oracle_slic
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    k = fn(lab.ctypes.data, rows, cols, int(step), int(nc), int(iterations), labels.ctypes.data, centers.ctypes.data)
    return labels, centers[:k].copy()
