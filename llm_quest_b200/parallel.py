"""Sample sharding across the GPUs of one box (one process per GPU, torch.distributed).

The vision-encode-and-fuse path has no cross-sample operation: rank r encodes
samples [r*B/G, (r+1)*B/G) with replicated weights and no communication. The only collective is the
all-gather (dim 0) of the fused embeddings — and of the [3, b, seq] position ids, gathered on
dim 1 — for a consumer that wants the whole batch on every rank. NCCL on GPUs, gloo in CPU tests.
"""

from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_samples: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced split: the first n % world ranks get one extra sample."""
    base, extra = divmod(n_samples, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(t: torch.Tensor, rank: int | None = None, world: int | None = None, dim: int = 0) -> torch.Tensor:
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_bounds(t.shape[dim], rank, world)
    return t.narrow(dim, lo, hi - lo)


def all_gather_cat(x: torch.Tensor, dim: int = 0, group=None) -> torch.Tensor:
    """Concatenate equally-shaped per-rank tensors along `dim` on every rank."""
    world = dist.get_world_size(group)
    if world == 1:
        return x
    xc = x.movedim(dim, 0).contiguous()
    out = torch.empty((world * xc.shape[0], *xc.shape[1:]), dtype=xc.dtype, device=xc.device)
    try:
        dist.all_gather_into_tensor(out, xc, group=group)
    except (RuntimeError, NotImplementedError):  # backends without the flat variant
        parts = list(out.chunk(world, dim=0))
        dist.all_gather(parts, xc, group=group)
    return out.movedim(0, dim)


def all_gather_fused(inputs_embs: torch.Tensor, position_ids: torch.Tensor | None = None, group=None):
    """inputs_embs [b_loc, seq, D] -> [b, seq, D]; position_ids [3, b_loc, seq] -> [3, b, seq]."""
    embs = all_gather_cat(inputs_embs, 0, group)
    pids = None if position_ids is None else all_gather_cat(position_ids, 1, group)
    return embs, pids
