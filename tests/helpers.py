"""Backend adapters for the parity tests: the same test body drives the product library on the GPU
(host-pointer and device-pointer entry points) or the CPU emulator build of the same kernels."""
from __future__ import annotations

import numpy as np

from depth_completion_mt_b200 import api


class Backend:
    def __init__(self, lib, mode: str):
        self.lib = lib
        self.mode = mode  # "emu" | "gpu_host" | "gpu_device"

    def _to(self, a):
        if self.mode != "gpu_device" or a is None:
            return a
        import torch

        return torch.from_numpy(np.ascontiguousarray(a)).cuda()

    @staticmethod
    def _from(x):
        if isinstance(x, tuple):
            return tuple(Backend._from(v) for v in x)
        if isinstance(x, np.ndarray) or x is None:
            return x
        return x.cpu().numpy()

    def img_completion(self, sparse, blur_type="gaussian", **kw):
        return self._from(api.img_completion(self._to(sparse), False, blur_type, lib=self.lib, **kw))

    def interpolate_with_superpixels(self, labels, sparse, use_superpixel=1, **kw):
        return self._from(api.interpolate_with_superpixels(self._to(labels), self._to(sparse), "gaussian", use_superpixel,
                                                           lib=self.lib, **kw))

    def stereo_refine(self, depth_ig, left, right, params=None, **kw):
        return self._from(api.stereo_refine(self._to(depth_ig), self._to(left), self._to(right), params, lib=self.lib, **kw))
