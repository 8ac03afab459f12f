"""Frame sharding across ranks: contiguous ceil(n/G) ranges, no collective on the data path.  The N > 1
host logic is exercised with world_size-2 gloo processes on the CPU (the kernels themselves need a GPU)."""
import os
import socket

import numpy as np
import pytest


def test_shard_ranges_partition_the_batch(tod):
    from tod_b200 import shard
    for n in (1, 2, 7, 64, 65, 4096):
        for g in (1, 2, 4, 8):
            ranges = [shard.shard_range(n, g, r) for r in range(g)]
            covered = [i for lo, hi in ranges for i in range(lo, hi)]
            assert covered == list(range(n))
            assert max(hi - lo for lo, hi in ranges) == -(-n // g)
    with pytest.raises(ValueError):
        shard.shard_range(4, 2, 2)


def _worker(rank, world, port, n, q):
    import torch.distributed as dist
    import tod_b200
    from tod_b200 import shard
    from tests import dist_helpers
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames = np.arange(n * 5, dtype=np.int64).reshape(n, 5)          # stand-in for per-frame results
    lo, hi = shard.shard_range(n, world, rank)
    mine = frames[lo:hi] * 2 + 1                                       # "process" this rank's shard
    gathered = dist_helpers.gather_frames(mine, n, world, rank)               # host-side gather, rank 0 only
    ms = dist_helpers.max_over_ranks(10.0 + rank)
    same = dist_helpers.all_equal_over_ranks(1234567) and not dist_helpers.all_equal_over_ranks(1000 + rank)                             # bench.py's max-over-ranks timing
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, None if gathered is None else gathered.tolist(), ms, same))


def test_two_rank_gloo_gather_is_order_independent():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n = 7
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    want = (np.arange(n * 5, dtype=np.int64).reshape(n, 5) * 2 + 1).tolist()
    assert res[0][1] == want and res[1][1] is None
    assert res[0][2] == res[1][2] == 11.0
    assert res[0][3] and res[1][3]
