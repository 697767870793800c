"""Data parallelism over LP instances.

LP instances are independent, so a minibatch is sharded across ranks in contiguous blocks with no
data-path collective; the only exchange is the all-reduce of the predictor's gradients after
`loss.backward()` (NCCL over NVLink on the GPU box, gloo in the CPU tests).  The reference has no
distributed code at all (it solves the 24 instances of a minibatch in a Python loop,
Interior/intopt_energy_mlp.py:409-453); summing the per-rank gradients reproduces exactly that loop's
gradient accumulation before `optimizer.step()` (:453).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, rank: int, world: int):
    """Contiguous block [lo, hi) of `n_items` owned by `rank`; blocks differ by at most one item."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(tensor: torch.Tensor, rank: int, world: int, dim: int = 0):
    lo, hi = shard_bounds(tensor.shape[dim], rank, world)
    return tensor.narrow(dim, lo, hi - lo)


def allreduce_gradients(parameters, group=None, average: bool = False):
    """Sum (or average) the .grad of `parameters` over the ranks in ONE flat all-reduce.

    The predictors are tiny (9 .. ~115k parameters), so a single latency-bound collective per step
    is the right granularity on NVSwitch; no bucketing, no overlap machinery."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    # EVERY parameter takes part, with zeros where a rank has no gradient (an empty shard, an unused head): the flat
    # buffers must have the same layout on every rank or the collective hangs / mixes parameters up
    params = [p for p in parameters if p.requires_grad]
    if not params:
        return
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for p in params:
        n = p.numel()
        piece = flat[off:off + n].view_as(p)
        if p.grad is None:
            p.grad = piece.clone()
        else:
            p.grad.copy_(piece)
        off += n


def broadcast_parameters(parameters, src: int = 0, group=None):
    """Make every rank start from rank `src`'s parameters (data-parallel steps only keep replicas equal if they start
    equal).  No-op without an initialised process group."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    with torch.no_grad():
        for p in parameters:
            dist.broadcast(p.data, src=src, group=group)


def zero_loss(parameters):
    """A loss of exactly zero that still touches every parameter: what a rank with an EMPTY shard back-propagates, so
    that its gradients exist (as zeros) and the all-reduce sees the same layout everywhere."""
    total = None
    for p in parameters:
        if p.requires_grad:
            t = p.sum() * 0.0
            total = t if total is None else total + t
    return total
