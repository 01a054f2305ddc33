"""Batch sharding of the detect path over the GPUs of one box (BASELINE config #4):
contiguous image ranges per rank, one process per GPU, NO collective on the data path.  The only
communication is bookkeeping for a sharded run: a barrier, the max-over-ranks of the device time and
the gather of per-rank result counts (torch.distributed: NCCL on GPUs, gloo in the CPU tests)."""


def shard_range(n_items, rank, world):
    """Contiguous [start, stop) of `n_items` for `rank` of `world`; sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world %r/%r" % (rank, world))
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def batches(start, stop, batch_size):
    """[(first, count)] batches of one shard; the last one may be short (caller pads it)."""
    return [(i, min(batch_size, stop - i)) for i in range(start, stop, batch_size)]


def max_over_ranks(value, device=None):
    """max of a python float over all ranks (identity without an initialised process group)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_counts(count, device=None):
    """[count of rank 0, count of rank 1, ...] on every rank."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [int(count)]
    t = torch.tensor([int(count)], dtype=torch.int64, device=device)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [int(o.item()) for o in out]
