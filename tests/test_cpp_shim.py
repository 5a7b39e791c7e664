"""include/img_completion.h (the header the reference includes but never shipped) compiles without OpenCV and a C++
caller written like src/DC_lidar_only/main.cpp:93 gets the oracle's bytes.  CPU: linked against the emulator
build; GPU: against libdcmt.so.  Without a GPU the product library makes the call fail loudly."""
from __future__ import annotations

import os
import subprocess

import numpy as np
import pytest

from depth_completion_mt_b200 import build, synth
from oracle import c_oracle as co
from tests.conftest import ROOT, assert_bit_equal


def compile_shim(tmp_path, lib_path):
    exe = str(tmp_path / "shim_main")
    libdir, libname = os.path.dirname(lib_path), os.path.basename(lib_path)[3:-3]
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "shim_main.cpp"), "-o", exe, "-L", libdir, "-l" + libname, "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def run_shim(exe, tmp_path, s, blur):
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    s.tofile(fin)
    r = subprocess.run([exe, str(s.shape[0]), str(s.shape[1]), fin, fout, blur], capture_output=True, text=True, timeout=600)
    return r, (np.fromfile(fout, np.float32).reshape(s.shape) if r.returncode == 0 else None)


def check_eval(stdout, gt, dense):
    """the shim's evaluate_performance / evaluate_performances against the literal float32 loops"""
    mse, rmse, mae, mae2, rmse2 = (float(v) for v in stdout.split())
    assert abs(mse - co.evaluate(gt, dense, 0, 0)["mean_err"]) < 1e-4
    ref = co.evaluate(gt, dense, 0, 1)
    assert abs(rmse - ref["rmse"]) < 1e-3 and abs(mae - ref["mae"]) < 1e-3
    ref = co.evaluate(gt, dense, 2, 1)
    assert abs(rmse2 - ref["rmse"]) < 1e-3 and abs(mae2 - ref["mae"]) < 1e-3


def test_shim_on_emulator_matches_oracle(tmp_path, emu_lib):
    exe = compile_shim(tmp_path, emu_lib.path)
    for blur in ("gaussian", "none", "something else"):
        s = synth.sparse_depth(17, 60, 100, 0.05)
        r, out = run_shim(exe, tmp_path, s, blur)
        assert r.returncode == 0, r.stderr
        assert_bit_equal(out, co.img_completion(s, blur if blur in ("gaussian", "none") else "none"), f"shim {blur}")
        check_eval(r.stdout, s, out)


def test_shim_with_product_library_fails_loudly_without_gpu(tmp_path):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    exe = compile_shim(tmp_path, build.build_library())
    r, _ = run_shim(exe, tmp_path, synth.sparse_depth(17, 40, 48, 0.05), "gaussian")
    assert r.returncode == 4 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_shim_with_product_library_on_gpu(tmp_path, gpu_lib):
    exe = compile_shim(tmp_path, gpu_lib.path)
    s = synth.sparse_depth(3)
    r, out = run_shim(exe, tmp_path, s, "gaussian")
    assert r.returncode == 0, r.stderr
    assert_bit_equal(out, co.img_completion(s, "gaussian"), "shim on the GPU")
    check_eval(r.stdout, s, out)
