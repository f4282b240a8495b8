"""Multi-GPU sharding of a batch: contiguous key ranges, no collective on the data path.

Every key / field element is independent, so rank g of W simply owns
elements [g*n/W, (g+1)*n/W).  The only optional exchange is gathering the result byte strings,
which is timed and reported separately from the ladder.
"""
from __future__ import annotations


def key_range(rank: int, world: int, n: int):
    """Half-open [lo, hi) of the n elements owned by `rank`; ranges are contiguous, disjoint,
    cover [0, n) and differ in size by at most one."""
    if not (0 <= rank < world) or n < 0:
        raise ValueError("bad shard request rank=%d world=%d n=%d" % (rank, world, n))
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_results(local, n_total: int, group=None):
    """All-gather per-rank result rows ([n_local, Nbytes] uint8 tensors) into key order.
    Works with NCCL (cuda tensors, NVLink) and gloo (cpu tensors, used by the CPU tests)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = [key_range(r, world, n_total) for r in range(world)]
    maxn = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((maxn,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([bufs[r][: hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)
