"""Multi-GPU layout: environments are independent, so the batch is split into contiguous ranges, one
process per GPU, with NO collective on the step path.  The only reductions are the
max-over-ranks of device timings / episode metrics."""
import numpy as np


def shard_envs(total_envs, rank, world_size, base_seed=1234):
    """(first env, one-past-last env, int64 seeds) owned by `rank`; seeds follow the reference's
    convention of consecutive integers (> 100 for training, help_initialization.py:194-201)."""
    per = (total_envs + world_size - 1) // world_size
    lo = min(total_envs, rank * per)
    hi = min(total_envs, lo + per)
    return lo, hi, base_seed + np.arange(lo, hi, dtype=np.int64)


def reduce_max(value, dist=None, device=None):
    """max over ranks of a python float (device timings); identity without a process group."""
    if dist is None or not dist.is_initialized():
        return float(value)
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
