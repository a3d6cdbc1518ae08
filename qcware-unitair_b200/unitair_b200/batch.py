"""Batch-dimension data parallelism: independent states shard trivially, no communication on
the data path (BASELINE.json north_star).

The reference has no distributed code.  With one process per GPU (torch.distributed), every
rank keeps a slice of the leading batch dimension of the states (and of per-entry gate
parameters); gates shared by the whole batch are replicated.  Forward and backward then run
on each rank's slice with the single-GPU engine.  Only the gradients of SHARED parameters need
a reduction: one all_reduce at the end of backward (`all_reduce_gradients`), a few kilobytes.

    theta = torch.nn.Parameter(...)                    # replicated
    states = shard_batch(all_states)                   # this rank's rows
    loss = loss_fn(theta, states)                      # local
    loss.backward()
    all_reduce_gradients([theta])                      # sum over ranks, one collective
"""
from __future__ import annotations

from typing import Iterable, Optional, Sequence

import torch
import torch.distributed as dist


def batch_slice(batch: int, rank: Optional[int] = None, world: Optional[int] = None) -> slice:
    """Rows [lo, hi) of a batch of `batch` entries owned by `rank` of `world` (defaults: the
    default process group).  Contiguous chunks; the first batch % world ranks hold one extra
    row (torch.tensor_split's rule), so any batch size works."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside a world of {world}")
    base, extra = divmod(int(batch), world)
    lo = rank * base + min(rank, extra)
    return slice(lo, lo + base + (1 if rank < extra else 0))


def shard_batch(t: torch.Tensor, dim: int = 0, group=None) -> torch.Tensor:
    """This rank's contiguous slice of batch dimension `dim` of `t` (a view; no communication).
    Use it on the states and on every per-entry operator / parameter tensor alike."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    return t.narrow(dim, *_start_len(batch_slice(t.shape[dim], rank, world)))


def _start_len(s: slice):
    return s.start, s.stop - s.start


def all_reduce_gradients(params: Iterable[torch.Tensor], group=None, average: bool = False) -> None:
    """Sum (or average) the .grad of SHARED parameters over all ranks with ONE collective: the
    gradients are flattened into a single buffer, all-reduced and copied back.  Parameters whose
    .grad is None on this rank contribute zeros (a rank's batch slice may not touch them)."""
    params = [p for p in params]
    if not params or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    real_views = []
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
        g = p.grad
        real_views.append(torch.view_as_real(g).reshape(-1) if g.is_complex() else g.reshape(-1))
    dtype = torch.promote_types(*([v.dtype for v in real_views] * 2)[:2]) if len(real_views) > 1 else real_views[0].dtype
    for v in real_views[2:]:
        dtype = torch.promote_types(dtype, v.dtype)
    flat = torch.cat([v.to(dtype) for v in real_views])
    dist.all_reduce(flat, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for v in real_views:
        n = v.numel()
        v.copy_(flat[off:off + n].to(v.dtype))
        off += n


def gather_batch(t: torch.Tensor, batch: int, dim: int = 0, group=None) -> torch.Tensor:
    """All ranks' slices of a batch-sharded result concatenated along `dim` (every rank gets the
    whole tensor; for small per-entry results such as expectation values)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return t
    world = dist.get_world_size(group)
    sizes = [_start_len(batch_slice(batch, r, world))[1] for r in range(world)]
    moved = t.movedim(dim, 0).contiguous()
    real = torch.view_as_real(moved) if moved.is_complex() else moved
    # all_gather wants equal sizes: pad every slice to the largest one, trim after
    big = max(sizes)
    padded = real
    if real.shape[0] < big:
        padded = torch.cat([real, real.new_zeros((big - real.shape[0],) + tuple(real.shape[1:]))], 0)
    parts = [torch.empty_like(padded) for _ in sizes]
    dist.all_gather(parts, padded.contiguous(), group=group)
    out = torch.cat([p[:s] for p, s in zip(parts, sizes)], 0)
    if moved.is_complex():
        out = torch.view_as_complex(out)
    return out.movedim(0, dim)
