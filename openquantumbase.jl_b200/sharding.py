"""Batch sharding and the one collective of the path.

Every ρ / ψ / schedule is independent inside an RHS call, so the batch is split into contiguous
blocks, one per rank (one process per GPU), with no data-path communication.  The only exchange
is the final ensemble average of observables: each rank reduces its members locally (`oqb_expect`
+ a sum), then one all-reduce of n_obs numbers follows (NCCL on GPUs, gloo in the CPU tests).
"""
from typing import Tuple


def shard_bounds(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of `total` batch members owned by `rank` (sizes differ by ≤ 1)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def ensemble_average(local_sum, local_count: int, group=None):
    """Global mean of per-member observables from per-rank partial sums (torch tensor, any device).
    Returns a new tensor; a single all_reduce(sum) carries [sums…, count]."""
    import torch
    import torch.distributed as dist
    flat = torch.cat([local_sum.reshape(-1).to(torch.complex128),
                      torch.tensor([float(local_count)], dtype=torch.complex128, device=local_sum.device)])
    buf = torch.view_as_real(flat).contiguous()
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    flat = torch.view_as_complex(buf)
    return (flat[:-1] / flat[-1].real).reshape(local_sum.shape)
