"""tiny SLIC workload for profiling: python tools/slic_run.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from depth_completion_mt_b200 import _lib, api, synth

lib = _lib.load()
lab = torch.from_numpy(synth.lab_image(0)).cuda()
for _ in range(3):
    api.generate_superpixels(lab, 18, 50, lib=lib)
torch.cuda.synchronize()
