"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz from the cv2 transliteration
(oracle/cv2_oracle.py, OpenCV 4.13.0), i.e. from the very library the reference calls.

    python -m oracle.make_golden

The fixtures are small (a few hundred kB in total) and committed, so that the plain-C oracle and
the CUDA path stay pinned to OpenCV's arithmetic on machines where cv2 or /root/reference do not
exist (the GPU box).  Re-running must reproduce the committed files byte for byte.

Where the reference build exists (oracle/_ref/libdcmt_ref.so: the reference's own sources compiled
from /root/reference, see oracle/ref_oracle.py), every fixture the reference computes is checked
against ITS output before it is written: the vectors are outputs of the reference itself.
"""
from __future__ import annotations

import hashlib
import os

import numpy as np

from depth_completion_mt_b200 import synth
from oracle import cv2_oracle as cvo

try:
    from oracle import ref_oracle as ro

    HAVE_REF = ro.available()
except Exception:  # pragma: no cover
    HAVE_REF = False


def pinned(what, arr, ref_fn):
    """arr as computed by the transliteration; must equal the reference build's own output bit for bit."""
    if HAVE_REF:
        ref = np.asarray(ref_fn())
        assert ref.shape == arr.shape and np.array_equal(ref.view(np.uint32), np.asarray(arr).view(np.uint32)), f"{what}: differs from the reference build"
    return arr


OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def multipass_frame(rows=120, cols=64):
    """Valid pixels only in the top and bottom three rows: the 31x31 fill needs several passes
    (img_completion.cpp:146-166 loops > 1 time)."""
    rng = np.random.default_rng(77)
    d16 = np.zeros((rows, cols), np.uint16)
    d16[:3] = rng.integers(512, 20480, (3, cols), dtype=np.uint16)
    d16[-3:] = rng.integers(512, 20480, (3, cols), dtype=np.uint16)
    return d16.astype(np.float32) / np.float32(256)


def lidar_cases():
    cases = {
        "q8_48x64_p05": synth.sparse_depth(0, 48, 64, 0.05),
        "q8_33x40_p02_kitti": synth.sparse_depth(1, 33, 40, 0.02, kitti_like=True),
        "q8_64x96_p20": synth.sparse_depth(2, 64, 96, 0.20),
        "q8_70x130_p01": synth.sparse_depth(3, 70, 130, 0.01),
        "q8_empty_20x30": np.zeros((20, 30), np.float32),
        "q8_1x1": synth.sparse_depth(4, 1, 1, 1.0),
        "q8_1x7": synth.sparse_depth(5, 1, 7, 0.5),
        "q8_7x1": synth.sparse_depth(6, 7, 1, 0.5),
        "q8_2x2": synth.sparse_depth(7, 2, 2, 0.5),
        "q8_3x5": synth.sparse_depth(8, 3, 5, 0.4),
        "multipass_120x64": multipass_frame(),
        "float_40x56_p05": synth.sparse_depth_float(0, 40, 56, 0.05),
        "float_neg_24x36": (synth.sparse_depth_float(1, 24, 36, 0.3, hi=110.0) - np.float32(3.0)).astype(np.float32),
    }
    return cases


def main():
    os.makedirs(OUT, exist_ok=True)
    # --- lidar only (img_completion.cpp:17-204)
    arrs = {}
    for name, s in lidar_cases().items():
        arrs[name + "__in"] = s
        for bt in ("gaussian", "none", "bilateral"):
            st = {}
            o = cvo.img_completion(s, bt, st)
            arrs[f"{name}__{bt}"] = o if bt == "bilateral" else pinned(f"{name} {bt}", o, lambda: ro.img_completion(s, bt))  # the reference's bilateral branch throws
            arrs[f"{name}__passes"] = np.int32(st["loop_passes"])
    np.savez_compressed(os.path.join(OUT, "lidar_only.npz"), **arrs)

    # --- single operators as OpenCV computes them (pin the C restatement operator by operator)
    rng = np.random.default_rng(5)
    x = (rng.random((37, 45)) * 100 - 5).astype(np.float32)
    import cv2
    ops = {"x": x, "two_tap": cv2.dilate(x, cvo.diamond_kernel_as_read_by_opencv()),
           "median5": cv2.medianBlur(x, 5), "gaussian5": cv2.GaussianBlur(x, (5, 5), 0),
           "bilateral5": cv2.bilateralFilter(x, 5, 1.5, 2.0)}
    for k in (5, 7, 31):
        ops[f"dilate{k}"] = cv2.dilate(x, np.ones((k, k), np.uint8))
        ops[f"erode{k}"] = cv2.erode(x, np.ones((k, k), np.uint8))
    ops["close5"] = cv2.morphologyEx(x, cv2.MORPH_CLOSE, np.ones((5, 5), np.uint8))
    xq = (rng.integers(0, 25600, (37, 45)).astype(np.float32) / np.float32(256))
    ops["xq"] = xq
    ops["gaussian5_q8"] = cv2.GaussianBlur(xq, (5, 5), 0)
    np.savez_compressed(os.path.join(OUT, "operators.npz"), **ops)

    # --- superpixel guided (img_completion_lc.cpp:34-203)
    arrs = {}
    for name, (r, c, p, step) in {"g_48x64_s9": (48, 64, 0.08, 9), "g_40x33_s6": (40, 33, 0.1, 6),
                                  "g_96x160_s18": (96, 160, 0.05, 18), "g_5x7_s2": (5, 7, 0.5, 2)}.items():
        s = synth.sparse_depth(11, r, c, p)
        lab, k = synth.superpixel_labels(11, r, c, step)
        lab[::7, ::5] = -1  # unassigned pixels (slic.cpp initialises clusters to -1)
        arrs[name + "__in"] = s
        arrs[name + "__labels"] = lab
        arrs[name + "__k"] = np.int32(k)
        for sp in (1, 0):
            arrs[name + f"__sp{sp}"] = pinned(f"{name} sp={sp}", cvo.interpolate_with_superpixels(s, lab, k, use_superpixel=sp),
                                              lambda: ro.interpolate_with_superpixels(lab, s, use_superpixel=sp, n_clusters=k))
    np.savez_compressed(os.path.join(OUT, "guided.npz"), **arrs)

    # --- stereo refinement (main_sl.cpp:715-885,1253); numpy float32 restatement + cv2 Gaussian
    arrs = {}
    for name, (r, c) in {"s_24x48": (24, 48), "s_40x200": (40, 200)}.items():
        dig, left, right = synth.stereo_pair(3, r, c)
        arrs[name + "__depth_ig"] = dig
        arrs[name + "__left"] = left
        arrs[name + "__right"] = right
        arrs[name + "__default"] = pinned(f"{name} stereo", cvo.stereo_refine(dig, left, right), lambda: ro.stereo_refine(dig, left, right))
        arrs[name + "__default_nogauss"] = pinned(f"{name} stereo, no blur", cvo.stereo_refine(dig, left, right, final_gauss=False),
                                                  lambda: ro.stereo_refine(dig, left, right, final_gauss=False))
        arrs[name + "__official10"] = cvo.stereo_refine(dig, left, right, num_iterations=10, damp_factor=1370.0,
                                                        err_clip=221.0, depth_clip=80.0, final_gauss=False)
        d0 = cvo.get_initial_disparity(dig)
        arrs[name + "__disp0"] = d0
        arrs[name + "__disp4"] = pinned(f"{name} optimize_IG", cvo.optimize_IG(left.astype(np.float32), right.astype(np.float32), d0),
                                        lambda: ro.optimize_IG(left.astype(np.float32), right.astype(np.float32), d0))
        dx, dy = cvo.measurement_derivatives(right.astype(np.float32))
        arrs[name + "__dx_right"] = dx
        arrs[name + "__dy_right"] = dy
    np.savez_compressed(os.path.join(OUT, "stereo.npz"), **arrs)

    # --- full KITTI size: digests only (inputs are re-generated from the seed by synth)
    lines = []
    for f in (0, 1, 2):
        for kitti_like in (False, True):
            s = synth.sparse_depth(f, density=0.05, kitti_like=kitti_like)
            for bt in ("gaussian", "none"):
                o = pinned(f"352x1216 frame {f} {bt}", cvo.img_completion(s, bt), lambda: ro.img_completion(s, bt))
                lines.append(f"{f} {int(kitti_like)} {bt} {hashlib.sha256(s.tobytes()).hexdigest()} {hashlib.sha256(o.tobytes()).hexdigest()}")
    with open(os.path.join(OUT, "lidar_only_352x1216.sha256"), "w") as fh:
        fh.write("# frame kitti_like blur sha256(input f32 bytes) sha256(output f32 bytes); synth.sparse_depth(frame, density=0.05)\n")
        fh.write("\n".join(lines) + "\n")
    print("golden written to", OUT, "(checked against the reference build)" if HAVE_REF else "(reference build not available: unchecked)")


if __name__ == "__main__":
    main()
