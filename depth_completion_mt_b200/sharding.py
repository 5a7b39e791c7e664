"""Frame-wise data parallelism: frames are independent, so a batch is partitioned
across ranks with NO collective on the hot path.  ``torch.distributed`` (NCCL on GPUs, gloo in the
CPU tests) only gathers per-rank timings and output checksums for validation and reporting.
One process per GPU, launched by torchrun."""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist


def env_rank_world() -> tuple[int, int, int]:
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def shard_range(n_frames: int, rank: int, world: int) -> tuple[int, int]:
    """Rank r of R processes frames [r*N/R, (r+1)*N/R) (balanced to within one frame)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return (n_frames * rank) // world, (n_frames * (rank + 1)) // world


def init_process_group(backend: str | None = None) -> tuple[int, int, int]:
    rank, local_rank, world = env_rank_world()
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, local_rank, world


def frame_checksums(frames) -> torch.Tensor:
    """Order-independent-per-frame 64-bit checksum of the raw float32 bits of each frame: sum of the
    int32 views with a position weight, kept in int64 (exact, identical on CPU and GPU)."""
    t = frames if isinstance(frames, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(frames))
    bits = t.contiguous().view(torch.int32).reshape(t.shape[0], -1).to(torch.int64)
    w = (torch.arange(bits.shape[1], device=bits.device, dtype=torch.int64) % 8191) + 1
    return (bits * w).sum(dim=1)


def gather_validation(elapsed_ms: float, frames_done: int, checksums: torch.Tensor, device=None) -> dict:
    """all_gather of {elapsed ms, frames, per-frame checksums}; every rank returns the same dict.
    ``max_ms`` (max over ranks) is the time throughput is computed from."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    device = device or checksums.device
    head = torch.tensor([elapsed_ms, float(frames_done)], dtype=torch.float64, device=device)
    if world == 1:
        heads, sums = [head], [checksums.to(device)]
    else:
        heads = [torch.empty_like(head) for _ in range(world)]
        dist.all_gather(heads, head)
        n = torch.tensor([checksums.numel()], dtype=torch.int64, device=device)
        ns = [torch.empty_like(n) for _ in range(world)]
        dist.all_gather(ns, n)
        m = int(max(int(x.item()) for x in ns))
        pad = torch.zeros(m, dtype=torch.int64, device=device)
        pad[: checksums.numel()] = checksums.to(device)
        sums_p = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(sums_p, pad)
        sums = [s[: int(k.item())] for s, k in zip(sums_p, ns)]
    ms = [float(h[0].item()) for h in heads]
    fr = [int(h[1].item()) for h in heads]
    return {"world": world, "ms_per_rank": ms, "frames_per_rank": fr, "max_ms": max(ms), "total_frames": sum(fr),
            "checksums": [s.cpu() for s in sums]}
