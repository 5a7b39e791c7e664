#!/usr/bin/env python
"""Throughput of the SURVEY.md 8(f) rows (SLIC, LiDAR projection, evaluation) at 352x1216 on one GPU, next to the C
oracle (the literal CPU loops, one core) on the same inputs.  One JSON line per row.

    python tools/bench_rows.py [--reps 20]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from depth_completion_mt_b200 import _lib, api, synth
from oracle import c_oracle as co


def gpu_time(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def cpu_time(fn, reps=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return 1e3 * (time.perf_counter() - t0) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    lib = _lib.load()
    rows, cols = 352, 1216
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    out = []
    # ---- evaluation: batch of 256 frames, 8 B/px
    n = 256
    gt = torch.from_numpy(np.stack([synth.sparse_depth(f, rows, cols, 0.3) for f in range(8)])).cuda().repeat(n // 8, 1, 1).contiguous()
    dense = (gt * 1.01 + 0.5).contiguous()
    ms = gpu_time(lambda: api.evaluate(gt, dense, "lidar_camera", lib=lib), a.reps)
    g0, d0 = gt[0].cpu().numpy(), dense[0].cpu().numpy()
    cms = cpu_time(lambda: co.evaluate(g0, d0, 0, 1))
    out.append({"row": "8f#3 evaluate_performance (main_lc.cpp:85-116)", "frames": n, "ms": ms, "frames_per_s": n / ms * 1e3,
                "achieved_GBps": 8 * rows * cols * n / ms / 1e6, "frac_of_hbm_peak": 8 * rows * cols * n / ms / 1e6 / peak,
                "cpu_oracle_ms_per_frame": cms, "note": "includes the read-back of 256 result records"})
    # ---- projection: one 120k-point cloud per call
    pts = torch.from_numpy(synth.velodyne_cloud(0, 120000)).cuda()
    ms = gpu_time(lambda: api.lidar_project(pts, synth.KITTI_T_VELO_TO_CAM, synth.KITTI_P_RECT_02, rows, cols, lib=lib), a.reps)
    pn = pts.cpu().numpy()
    cms = cpu_time(lambda: co.lidar_project(pn, synth.KITTI_T_VELO_TO_CAM, synth.KITTI_P_RECT_02, rows, cols))
    out.append({"row": "8f#2 LiDAR projection + normalize (main_sl.cpp:478-523)", "points": 120000, "ms": ms, "clouds_per_s": 1e3 / ms,
                "cpu_oracle_ms_per_cloud": cms, "note": "4 small kernels per cloud: launch-latency bound at this size"})
    # ---- projection, batched: 128 clouds per call in the same four launches
    nb = 128
    clouds = torch.from_numpy(np.stack([synth.velodyne_cloud(f % 8, 120000) for f in range(nb)])).cuda()
    ms = gpu_time(lambda: api.lidar_project_batch(clouds, None, synth.KITTI_T_VELO_TO_CAM, synth.KITTI_P_RECT_02, rows, cols, lib=lib), max(3, a.reps // 2))
    out.append({"row": "8f#2 LiDAR projection + normalize, batch of 128 clouds (dcmt_lidar_project_batch_f32)", "points": 120000, "ms": ms,
                "clouds_per_s": nb * 1e3 / ms})
    del clouds
    # ---- SLIC: one frame, step 18, 10 iterations
    lab = torch.from_numpy(synth.lab_image(0, rows, cols)).cuda()
    ms = gpu_time(lambda: api.generate_superpixels(lab, 18, 50, lib=lib), a.reps)
    ln = lab.cpu().numpy()
    cms = cpu_time(lambda: co.slic(ln, 18, 50), reps=1)
    out.append({"row": "8f#1 Slic::generate_superpixels (slic.cpp:101-182)", "step": 18, "iterations": 10, "ms": ms, "frames_per_s": 1e3 / ms,
                "cpu_oracle_ms_per_frame": cms})
    labs = torch.from_numpy(np.stack([synth.lab_image(f, rows, cols) for f in range(8)])).cuda().repeat(8, 1, 1, 1).contiguous()
    ms = gpu_time(lambda: api.generate_superpixels(labs, 18, 50, lib=lib), max(3, a.reps // 4))
    out.append({"row": "8f#1 Slic::generate_superpixels, batch of 64 frames", "step": 18, "iterations": 10, "ms": ms, "frames_per_s": 64e3 / ms})
    labs256 = labs.repeat(4, 1, 1, 1).contiguous()
    ms = gpu_time(lambda: api.generate_superpixels(labs256, 18, 50, lib=lib), 3)
    out.append({"row": "8f#1 Slic::generate_superpixels, batch of 256 frames", "step": 18, "iterations": 10, "ms": ms, "frames_per_s": 256e3 / ms})
    del labs256
    # ---- the DC_lidar_camera chain on device: SLIC -> guided completion -> evaluation, one frame (main_lc.cpp:184-225)
    sparse = torch.from_numpy(synth.sparse_depth(0, rows, cols, 0.05)).cuda()
    k = lib.dcmt_slic_center_count(rows, cols, 18)

    def chain():
        labels = api.generate_superpixels(lab, 18, 50, lib=lib)
        d = api.interpolate_with_superpixels(labels, sparse, "gaussian", 1, n_clusters=k, lib=lib)
        return api.evaluate(sparse, d, "lidar_camera", lib=lib)
    ms = gpu_time(chain, max(3, a.reps // 4))
    out.append({"row": "DC_lidar_camera chain: SLIC -> interpolate_with_superpixels -> evaluate_performance (main_lc.cpp:184-225)", "ms": ms,
                "frames_per_s": 1e3 / ms})
    sparse64 = torch.from_numpy(np.stack([synth.sparse_depth(f, rows, cols, 0.05) for f in range(8)])).cuda().repeat(8, 1, 1).contiguous()

    def chain64():
        labels = api.generate_superpixels(labs, 18, 50, lib=lib)
        d = api.interpolate_with_superpixels(labels, sparse64, "gaussian", 1, n_clusters=k, lib=lib)
        return api.evaluate(sparse64, d, "lidar_camera", lib=lib)
    ms = gpu_time(chain64, 3)
    out.append({"row": "DC_lidar_camera chain, batch of 64 frames", "ms": ms, "frames_per_s": 64e3 / ms})
    for o in out:
        print(json.dumps(o), flush=True)


if __name__ == "__main__":
    main()
