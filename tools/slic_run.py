"""SLIC workload for profiling: python tools/slic_run.py [frames]  (step 18, nc 50, 10 iterations: main_lc.cpp:187-201)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from depth_completion_mt_b200 import _lib, api, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
lib = _lib.load()
lab = torch.from_numpy(np.stack([synth.lab_image(f % 8) for f in range(n)])).cuda()
for _ in range(3):
    api.generate_superpixels(lab, 18, 50, lib=lib)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    api.generate_superpixels(lab, 18, 50, lib=lib)
torch.cuda.synchronize()
print(f"SLIC {n} frames: {n * 5 / (time.perf_counter() - t0):.0f} frames/s")
