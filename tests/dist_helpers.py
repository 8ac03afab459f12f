"""Host-side multi-rank helpers shared by bench.py and tests/test_sharding.py (torch.distributed is plumbing for the
benchmark contract only; the product package never imports it)."""
import numpy as np


def gather_frames(mine, n, world, rank):
    """Collect every rank's per-frame results on rank 0 in frame order (torch.distributed must be initialised
    when world > 1).  Returns the [n, ...] array on rank 0 and None elsewhere."""
    if world == 1:
        return np.asarray(mine)
    import torch.distributed as dist
    parts = [None] * world if rank == 0 else None
    dist.gather_object(np.asarray(mine), parts, dst=0)
    if rank != 0:
        return None
    out = np.concatenate([p for p in parts if len(p)], axis=0)
    assert out.shape[0] == n
    return out


def max_over_ranks(value):
    """Timing rule of bench.py: a multi-GPU number is the slowest rank's."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def all_equal_over_ranks(value):
    """True on every rank iff all ranks hold the same integer (bench.py: output CRCs do not depend on the rank / GPU)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return True
    t = torch.tensor([int(value), -int(value)], dtype=torch.int64)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(t[0].item()) == -int(t[1].item())
