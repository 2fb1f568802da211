"""Frame sharding for the multi-GPU batch path.

Frames are independent in ICER (context model and entropy coder reset per segment-plane, no inter-frame state), so a
batch shards by frame index with no data-path collective: rank r of `world` codes frames [lo, hi).  The only
communication is bookkeeping: a barrier around timed regions and a MAX reduction of per-rank device times (and,
optionally, a gather of the per-frame compressed sizes).  Works with any torch.distributed backend (NCCL on the GPU box,
gloo in the CPU tests)."""


def shard_range(n, rank, world):
    """Contiguous, balanced split of n frames: the first n % world ranks get one extra frame."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_over_ranks(values, device=None):
    """Element-wise MAX of a list of floats over all ranks (timing: the slowest rank defines the step)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def gather_sizes(local_sizes, n_total, rank, world, device=None):
    """All ranks learn the compressed size of every frame of the global batch (frame order = shard order)."""
    import torch
    import torch.distributed as dist
    out = torch.zeros(n_total, dtype=torch.int64, device=device)
    lo, hi = shard_range(n_total, rank, world)
    out[lo:hi] = torch.as_tensor(local_sizes, dtype=torch.int64, device=device)
    if dist.is_available() and dist.is_initialized() and world > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM)
    return out
