"""Multi-rank plumbing (one process per GPU).  The only collective on the path is one sum all-reduce of the flat
gradient buffer per step; utterances are sharded by batch with no other exchange."""
import torch
import torch.distributed as dist


def world_size(group=None):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def rank(group=None):
    return dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0


def shard_rows(n, rnk=None, world=None):
    """Contiguous shard [lo, hi) of n utterances for this rank (sizes differ by at most one)."""
    rnk = rank() if rnk is None else rnk
    world = world_size() if world is None else world
    base, rem = divmod(n, world)
    lo = rnk * base + min(rnk, rem)
    return lo, lo + base + (1 if rnk < rem else 0)


def allreduce_sum_(flat, group=None):
    """In-place sum over ranks of the flat gradient buffer; returns the factor that turns the sum into the
    mean-over-ranks gradient (folded into the Adam kernel rather than applied as a separate pass)."""
    w = world_size(group)
    if w > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / w


def broadcast_from_rank0_(engine, group=None):
    """DDP's constructor-time broadcast: every rank takes rank 0's parameters and BatchNorm buffers (the reference never
    seeds the RNG, so without this each replica would start from its own random init and never agree)."""
    if world_size(group) <= 1:
        return
    engine.sync()
    dist.broadcast(engine.flat, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    if engine.flat_buf is not None and engine.flat_buf.numel():
        dist.broadcast(engine.flat_buf, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
