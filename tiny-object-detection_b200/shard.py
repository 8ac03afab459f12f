"""Frame sharding rule of the multi-GPU path: camera frames are independent, so a batch is cut into
contiguous ranges of ceil(n / G), one per GPU / rank, and nothing is exchanged on the data path.  The library applies the
same rule inside `tod_pool_*` (csrc/pool.cu::shard); this is its host-language mirror for callers that run one process
per GPU (bench.py under torchrun)."""


def shard_range(n, world, rank):
    """[lo, hi) of the frames rank `rank` processes: contiguous blocks of ceil(n / world)."""
    if not 0 <= rank < world:
        raise ValueError("rank %d outside [0, %d)" % (rank, world))
    per = -(-n // world)
    lo = min(n, rank * per)
    return lo, min(n, lo + per)
