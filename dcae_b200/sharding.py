"""Multi-GPU host logic of the slice loop: images are independent, so ranks take disjoint
image shards and the only data-path reduction is one scalar all-reduce for bpp (train.py:82-85).
Backend-agnostic (NCCL on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Sequence

import torch


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Rank r processes images r, r + world, r + 2 world, ... (batch sharding, weights replicated)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    return list(range(rank, n_items, world))


def reduce_bpp(log2_lik_sum: torch.Tensor, num_pixels: int, group=None) -> float:
    """bpp over all ranks = -sum_r(sum log2 lik) / sum_r(pixels): one all-reduce of two doubles."""
    import torch.distributed as dist
    v = torch.stack([log2_lik_sum.detach().double().reshape(()),
                     torch.tensor(float(num_pixels), dtype=torch.double, device=log2_lik_sum.device)])
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(v, op=dist.ReduceOp.SUM, group=group)
    return float(-v[0] / v[1])


def max_over_ranks(value_ms: float, device, group=None) -> float:
    """Multi-GPU timings are the max over ranks of the device time."""
    import torch.distributed as dist
    t = torch.tensor([float(value_ms)], dtype=torch.double, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t)


def gather_symbol_streams(per_rank: Sequence[torch.Tensor], order: Sequence[Sequence[int]]) -> torch.Tensor:
    """Re-interleave per-rank [5, B_r, 64, h, w] symbol (or index) tensors into global image order."""
    n = sum(len(o) for o in order)
    first = per_rank[0]
    out = first.new_empty((first.shape[0], n) + tuple(first.shape[2:]))
    for t, idxs in zip(per_rank, order):
        for j, g in enumerate(idxs):
            out[:, g] = t[:, j]
    return out
