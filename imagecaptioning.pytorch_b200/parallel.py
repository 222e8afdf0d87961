"""Data-parallel plumbing for the hot path: one process per GPU, torch.distributed (NCCL on GPUs, gloo in the CPU tests).

Images are independent, so
  * inference shards the image batch contiguously across ranks and needs NO data-path collective; ``gather_captions`` is the optional
    single all_gather of the [B/G, T] id matrices when one rank wants the whole result (tools/train_pl.py:224,269 gather pickled
    predictions instead);
  * SCST training keeps sampling / greedy baseline / reward / loss rank-local and issues exactly ONE all-reduce over a flat buffer of
    all gradients per step, averaging over ranks -- the arithmetic torch DDP's bucketed all-reduce performs for the reference
    (tools/train_pl.py:479), and what ``loss.mean()`` over DataParallel replicas amounts to (tools/train.py:188).
"""
from __future__ import annotations

from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, end) of ``n_items`` owned by ``rank``; earlier ranks take the remainder (sizes differ by at most 1)."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_batch(tensors: Iterable[torch.Tensor], rank: int, world: int) -> List[torch.Tensor]:
    tensors = list(tensors)
    s, e = shard_range(tensors[0].shape[0], rank, world)
    return [t[s:e] if t is not None else None for t in tensors]


def gather_captions(seq_local: torch.Tensor, n_total: int) -> torch.Tensor:
    """All ranks get the [n_total, T] id matrix in image order (ranks may own different row counts)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return seq_local
    world = dist.get_world_size()
    T = seq_local.shape[1]
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    max_rows = max(e - s for s, e in sizes)
    pad = torch.zeros(max_rows, T, dtype=seq_local.dtype, device=seq_local.device)
    pad[:seq_local.shape[0]] = seq_local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[:e - s] for o, (s, e) in zip(out, sizes)], 0)


def allreduce_gradients(params: Iterable[torch.nn.Parameter]) -> int:
    """One all-reduce(sum) over a single flat buffer holding every gradient, then divide by the world size.  Returns the byte count."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return 0
    flat = torch.cat([g.reshape(-1) for g in grads])
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(dist.get_world_size())
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return flat.numel() * flat.element_size()


def max_over_ranks(value: float, device) -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
