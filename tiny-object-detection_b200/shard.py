"""Frame sharding for the multi-GPU path: camera frames are independent, so a batch is cut into
contiguous ranges, one per rank / GPU, and nothing is exchanged on the data path.  The only cross-rank steps are
host-side: gathering per-frame results on rank 0 and the max-over-ranks of a timing."""
import numpy as np


def shard_range(n, world, rank):
    """[lo, hi) of the frames rank `rank` processes: contiguous blocks of ceil(n / world)."""
    if not 0 <= rank < world:
        raise ValueError("rank %d outside [0, %d)" % (rank, world))
    per = -(-n // world)
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


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
