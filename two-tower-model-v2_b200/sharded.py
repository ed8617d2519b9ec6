"""Catalog sharding across GPUs (north_star (3); the reference has no multi-device path).

One process per GPU.  Rank r owns the contiguous catalog rows [lo_r, hi_r); every rank searches
its shard for the replicated query batch, the per-rank (scores, ids) [nq,K] lists are exchanged
with ONE all-gather (NCCL over NVLink/NVSwitch), and a merge kernel (tt_topk_merge) reduces the G
sorted lists to the global top-K on every rank.  ids are global row numbers (local row + lo_r).

The local-search and merge steps are injectable so that the exchange logic is testable on CPU with
the gloo backend (tests/test_sharded_gloo.py).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_total: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block partition: rank r owns [r*ceil(N/G), min(N, (r+1)*ceil(N/G)))."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad world_size/rank")
    per = -(-n_total // world_size)
    lo = min(n_total, rank * per)
    hi = min(n_total, lo + per)
    return lo, hi


def gather_and_merge(scores: torch.Tensor, ids: torch.Tensor, merge: Callable, group=None):
    """All-gathers per-rank [nq,K] lists into [G,nq,K] and merges them."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return scores, ids
    nq, k = scores.shape
    sg = torch.empty((world, nq, k), device=scores.device, dtype=scores.dtype)
    ig = torch.empty((world, nq, k), device=ids.device, dtype=ids.dtype)
    dist.all_gather_into_tensor(sg.view(world * nq, k), scores.contiguous(), group=group)
    dist.all_gather_into_tensor(ig.view(world * nq, k), ids.contiguous(), group=group)
    return merge(sg, ig)


class ShardedFlatIPIndex:
    """A FlatIPIndex per rank over its row block + all-gather/merge search."""

    def __init__(self, local_index, n_total: int, group=None, merge: Optional[Callable] = None):
        self.local = local_index
        self.n_total = int(n_total)
        self.group = group
        if merge is None:
            from . import ops
            merge = ops.topk_merge
        self._merge = merge

    @property
    def ntotal(self) -> int:
        return self.n_total

    def search_device(self, q: torch.Tensor, k: int):
        """q replicated on every rank -> global (scores, ids) [nq,k] on every rank, and the number of
        local queries that needed the exact re-run."""
        k_local = min(k, self.local.ntotal)
        scores, ids, n_bad = self.local.search_checked_device(q, k_local)
        if k_local < k:   # shard smaller than k: pad so every rank contributes [nq,k]
            pad_s = torch.full((q.shape[0], k - k_local), float("-inf"), device=scores.device)
            pad_i = torch.full((q.shape[0], k - k_local), -1, device=ids.device, dtype=ids.dtype)
            scores, ids = torch.cat([scores, pad_s], 1), torch.cat([ids, pad_i], 1)
        s, i = gather_and_merge(scores, ids, self._merge, self.group)
        return s, i, n_bad
