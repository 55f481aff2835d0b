"""opencv_b200.batch -- sharding a batch of independent frames over the GPUs of one box (one process per GPU).

The hot path has no cross-frame state: rank r owns a contiguous block of frames, runs the whole per-frame
pipeline on its own stream(s) and keeps its results.  The ONLY exchange is a broadcast of the small shared operand
(template / filter taps / warp matrix) from rank 0 -- NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
import numpy as np


def shard_range(n_frames, rank, world):
    """contiguous block [lo, hi) of frames owned by `rank`; blocks differ in size by at most one frame"""
    base, rem = divmod(int(n_frames), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_operand(arr, src=0, device=None):
    """broadcast a small numpy operand (taps, template, matrix) from rank `src`; returns the numpy array on every rank.
    Works with any initialised torch.distributed backend; a no-op without one."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return np.asarray(arr)
    a = np.ascontiguousarray(arr)
    t = torch.from_numpy(a.copy())
    if device is not None:
        t = t.to(device)
    dist.broadcast(t, src)
    return t.cpu().numpy()


def gather_counts(local_count, device=None):
    """per-rank frame counts (for reporting whole-job throughput)"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [int(local_count)]
    t = torch.tensor([int(local_count)], dtype=torch.int64, device=device)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [int(x.item()) for x in out]
