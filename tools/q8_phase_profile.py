"""Per-phase cycle profile of the fused kernels from in-kernel clock64() stamps (dcmt_debug_q8_phase_cycles).
Run on the GPU box:  python tools/q8_phase_profile.py [frames]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from depth_completion_mt_b200 import _lib, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 79
lib = _lib.load()
frames = np.stack([synth.sparse_depth(f % 64) for f in range(n)])
d_in = torch.from_numpy(frames).cuda()
d_out = torch.empty_like(d_in)
rows, cols = frames.shape[1:]
max_tiles = ((rows + 31) // 32) * ((cols + 31) // 32)
fs = torch.zeros(n * max_tiles * 16, dtype=torch.int64, device="cuda")
ts = torch.zeros(n * max_tiles * 16, dtype=torch.int64, device="cuda")
tiles = C.c_int(0)
for _ in range(3):
    lib.check(lib.dcmt_debug_q8_phase_cycles(d_in.data_ptr(), d_out.data_ptr(), rows, cols, n, fs.data_ptr(), ts.data_ptr(), C.byref(tiles),
                                             int(torch.cuda.current_stream().cuda_stream)))
torch.cuda.synchronize()
nt = tiles.value * n
for name, buf, k, labels in (("k_q8_front", fs, 5, ["load+encode", "6 morphology passes", "final pass + store", "column keys"]),
                             ("k_q8_tail", ts, 10, ["load", "A5 extrapolation", "vertical 16-row maxima", "hole scan", "lazy 31-wide fill",
                                                    "replicate border", "median", "reflect border", "gaussian + store"])):
    a = buf.view(-1, 16)[:, :k].cpu().numpy().astype(np.float64)
    a = a[a[:, 0] != 0]  # the front runs on its own (lower) tiles: count the CTAs that left stamps
    nt = len(a)
    d = np.diff(a, axis=1)
    tot = a[:, -1] - a[:, 0]
    print(f"{name}: {nt} CTAs, mean CTA lifetime {tot.mean():.0f} cycles (min {tot.min():.0f}, max {tot.max():.0f})")
    for i, lab in enumerate(labels):
        print(f"   {lab:28s} {d[:, i].mean():9.0f} cycles  {d[:, i].mean() / tot.mean():6.1%}")
