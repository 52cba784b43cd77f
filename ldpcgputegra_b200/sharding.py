"""Frame sharding across GPUs: frames are independent, so rank r of W decodes a contiguous slice on its own GPU and only
three integers per rank — frames, bit errors, frame errors — are summed afterwards.  There is no collective on the decode path.

The reference has no multi-GPU code at all (`cudaSetDevice(0)` is hard-coded, code/gpu_fixed/main.cpp:113); its closest construct is one
decoder object per OpenMP section (code/gpu_fixed/test.cpp:241-281), which is what one process per GPU generalises.
"""
from __future__ import annotations


def shard(total_frames: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous partition: (first_frame, frame_count) of `rank`; counts differ by at most one and sum to the total."""
    if world < 1 or not 0 <= rank < world or total_frames < 0:
        raise ValueError("bad shard request")
    base, rem = divmod(total_frames, world)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def reduce_counters(counters, dist=None):
    """Sum [frames, bit_errors, frame_errors] over ranks.  `dist` is torch.distributed (any backend) or None for a single process."""
    vals = [int(c) for c in counters]
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return vals
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor(vals, dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [int(x) for x in t.tolist()]


def max_over_ranks(x: float, dist=None) -> float:
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
