"""TEST INFRASTRUCTURE ONLY: compiles the kernel sources of depth_completion_mt_b200/csrc with g++
against tests/emu/cuda_emu.h into tests/emu/libdcmt_emu.so -- the same C ABI as libdcmt.so, executed
by a CPU fiber emulator of the CUDA execution model.  Used by `pytest -m "not gpu"` to diff the
real kernel logic against the oracle where no GPU exists.  The product never loads this file."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "depth_completion_mt_b200", "csrc")
OUT = os.path.join(HERE, "libdcmt_emu.so")


def build(force: bool = False) -> str:
    srcs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(HERE, "cuda_emu.h"), os.path.join(HERE, "cuda_emu.cpp"), os.path.join(ROOT, "include", "dcmt.h")]
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= max(os.path.getmtime(d) for d in deps):
        return OUT
    obj_dir = os.path.join(HERE, "build")
    os.makedirs(obj_dir, exist_ok=True)
    flags = ["-O2", "-g", "-std=c++17", "-fPIC", "-DDCMT_EMU", "-ffp-contract=off", "-fvisibility=hidden",
             "-Wno-attributes", "-I", HERE, "-I", CSRC, "-I", os.path.join(ROOT, "include")]
    procs, objs = [], []
    for s in srcs + [os.path.join(HERE, "cuda_emu.cpp")]:
        o = os.path.join(obj_dir, os.path.basename(s).rsplit(".", 1)[0] + ".o")
        objs.append(o)
        cmd = ["g++"] + flags + ["-x", "c++", "-c", s, "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode:
            raise RuntimeError("emulator build failed: " + " ".join(cmd) + "\n" + out)
    r = subprocess.run(["g++", "-shared", "-o", OUT] + objs, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode:
        raise RuntimeError("emulator link failed\n" + r.stdout)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
