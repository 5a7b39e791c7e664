"""batched LiDAR projection workload for profiling: python tools/proj_run.py [clouds]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from depth_completion_mt_b200 import _lib, api, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
lib = _lib.load()
clouds = torch.from_numpy(np.stack([synth.velodyne_cloud(f % 8, 120000) for f in range(n)])).cuda()
for _ in range(4):
    api.lidar_project_batch(clouds, None, synth.KITTI_T_VELO_TO_CAM, synth.KITTI_P_RECT_02, 352, 1216, lib=lib)
torch.cuda.synchronize()
