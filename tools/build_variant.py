"""Build an A/B variant of libdcmt.so with extra nvcc flags: python tools/build_variant.py NAME -DDCMT_QTT=448 ...
-> depth_completion_mt_b200/variants/libdcmt_NAME.so (select with DCMT_LIB=<path>)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from depth_completion_mt_b200 import build as b  # noqa: E402

name, extra = sys.argv[1], sys.argv[2:]
vdir = os.path.join(b.PKG_DIR, "variants")
odir = os.path.join(vdir, "obj_" + name)
os.makedirs(odir, exist_ok=True)
objs, procs = [], []
for src in b.sources():
    obj = os.path.join(odir, os.path.basename(src)[:-3] + ".o")
    objs.append(obj)
    procs.append(subprocess.Popen([b._nvcc()] + b.NVCC_FLAGS + extra + ["-Xptxas", "-v", "-c", src, "-o", obj], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
for p in procs:
    out, _ = p.communicate()
    for l in out.splitlines():
        if "error" in l or ("k_q8_" in l and "Compiling" in l) or "spill" in l and "0 bytes spill" not in l:
            print(l[:200])
    assert p.returncode == 0, out
lib = os.path.join(vdir, f"libdcmt_{name}.so")
subprocess.run([b._nvcc()] + b.ARCH + ["-shared", "-o", lib] + objs, check=True)
print(lib)
