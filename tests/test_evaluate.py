"""Evaluation reductions against the literal float32 loops of the reference (oracle) and against exact
float64 sums.  Bar: counts equal; our double-accumulated sums within 1e-12 relative of numpy float64; the derived
mean_err / mae / rmse within 2e-4 relative (+1e-6 absolute) of the reference's float32 raster-order figures -- the
reference's own accumulation error at 4e5 pixels."""
from __future__ import annotations

import numpy as np
import pytest

from depth_completion_mt_b200 import api, synth
from oracle import c_oracle as co


def frames(seed, rows, cols):
    rng = np.random.default_rng(seed)
    gt = np.where(rng.random((rows, cols)) < 0.3, rng.uniform(0.5, 80, (rows, cols)), 0).astype(np.float32)
    dense = (rng.uniform(0.0, 85, (rows, cols)) * (rng.random((rows, cols)) < 0.9)).astype(np.float32)
    return gt, dense


def body(lib, to_backend):
    for rows, cols in ((37, 53), (352, 1216), (64, 100)):
        gt, dense = frames(rows, rows, cols)
        for variant, (mode, tol) in api.EVAL_VARIANTS.items():
            rec = api.evaluate(to_backend(gt), to_backend(dense), variant, lib=lib)[0]
            ref = co.evaluate(gt, dense, int(tol), mode)
            mask = (gt > tol) if mode == 0 else ((gt > tol) & (dense > tol))
            d = (gt - dense)[mask].astype(np.float64)  # float32 subtraction, then exact sums
            assert int(rec["count"]) == ref["count"] == int(mask.sum())
            for got, want in ((rec["sum_err"], d.sum()), (rec["sum_abs"], np.abs(d).sum()), (rec["sum_sq"], (d * d).sum())):
                assert abs(got - want) <= 1e-12 * max(1.0, abs(want)), (variant, got, want)
            for k in ("mean_err", "mae", "rmse"):
                if mode == 0 and k != "mean_err" or mode == 1 and k == "mean_err":
                    continue
                assert abs(float(rec[k]) - ref[k]) <= 2e-4 * abs(ref[k]) + 1e-6, (variant, k, float(rec[k]), ref[k])
    # the reference call surface
    gt, dense = frames(5, 48, 64)
    assert abs(api.evaluate_performance(to_backend(gt), to_backend(dense), "lidar_only", lib=lib) - co.evaluate(gt, dense, 0, 0)["mean_err"]) < 1e-4
    mse, mae = api.evaluate_performance(to_backend(gt), to_backend(dense), "lidar_camera", lib=lib)
    ref = co.evaluate(gt, dense, 0, 1)
    assert abs(mse - ref["rmse"]) < 1e-3 and abs(mae - ref["mae"]) < 1e-3
    mae2, rmse2 = api.evaluate_performances(to_backend(gt), to_backend(dense), lib=lib)
    ref2 = co.evaluate(gt, dense, 2, 1)
    assert abs(rmse2 - ref2["rmse"]) < 1e-3 and abs(mae2 - ref2["mae"]) < 1e-3
    # batch, deterministic, empty mask -> NaN like the reference's 0 / 0
    b_gt = np.stack([frames(s, 40, 72)[0] for s in range(4)] + [np.zeros((40, 72), np.float32)])
    b_r = np.stack([frames(s, 40, 72)[1] for s in range(5)])
    r1 = api.evaluate(to_backend(b_gt), to_backend(b_r), "stereo_lidar", lib=lib)
    r2 = api.evaluate(to_backend(b_gt), to_backend(b_r), "stereo_lidar", lib=lib)
    assert r1.tobytes() == r2.tobytes() or (np.isnan(r1["mae"][-1]) and r1[:4].tobytes() == r2[:4].tobytes())
    assert r1["count"][-1] == 0 and np.isnan(r1["mae"][-1]) and np.isnan(r1["rmse"][-1])
    for f in range(4):
        assert int(r1["count"][f]) == co.evaluate(b_gt[f], b_r[f], 2, 1)["count"]
    # completion output feeds evaluation (main.cpp:93-101)
    sparse = synth.sparse_depth(3, 64, 96, 0.05)
    dense = api.img_completion(to_backend(sparse), False, "gaussian", lib=lib)
    dn = dense if isinstance(dense, np.ndarray) else dense.cpu().numpy()
    got = api.evaluate_performance(to_backend(sparse), dense, "lidar_only", lib=lib)
    assert abs(got - co.evaluate(sparse, dn, 0, 0)["mean_err"]) < 1e-4


def test_emu_evaluate(emu_lib):
    body(emu_lib, lambda a: a)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["host", "device"])
def test_gpu_evaluate(gpu_lib, mode):
    import torch

    body(gpu_lib, (lambda a: a) if mode == "host" else (lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()))
