"""N > 1 host logic on CPU: world_size-2 gloo run of the sharding / timing helpers bench.py uses under torchrun."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from icer_compression_b200.shard import gather_sizes, max_over_ranks, shard_range


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 1024, 8191):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                lo, hi = shard_range(n, r, world)
                assert 0 <= lo <= hi <= n and hi - lo in (n // world, n // world + 1)
                got += list(range(lo, hi))
            assert got == list(range(n))
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 11
    lo, hi = shard_range(n, rank, world)
    sizes = [1000 * (f + 1) for f in range(lo, hi)]          # stand-in for size_used of this rank's frames
    allsizes = gather_sizes(sizes, n, rank, world)
    t = max_over_ranks([10.0 + rank, 5.0 - rank])
    dist.barrier()
    q.put((rank, allsizes.tolist(), t))
    dist.destroy_process_group()


def test_two_ranks_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, allsizes, t in res:
        assert allsizes == [1000 * (f + 1) for f in range(11)]
        assert t == [11.0, 5.0]
