"""Buyer tower: drop-in for the reference `src/models/buyer_tower.py:9-144`.

Same constructor, attributes (`embedding_dim`, `aggregation_method`, `attention`), state-dict
keys (`attention.0.weight`, `attention.0.bias`, `attention.2.weight`, `attention.2.bias`),
method names, argument meaning and exceptions.  The arithmetic runs in hand-written sm_100a CUDA
kernels (csrc/pool.cu, csrc/attn_logits.cu) through the C-ABI; there is no CPU path: CPU tensors
raise RuntimeError.

Beyond the reference surface: `precompute_item_logits` / `forward_gather` pool
straight out of a device-resident item-embedding table given history row indices, so a request
never re-encodes history items (reference: src/inference/encoder.py:276-303).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import ops


class BuyerTower(nn.Module):
    """Buyer Tower for encoding buyer behaviour into embeddings (reference buyer_tower.py:9)."""

    def __init__(self, embedding_dim: int = 384, aggregation_method: str = "attention",
                 attention_hidden_dim: int = 128):
        super().__init__()
        self.embedding_dim = embedding_dim
        self.aggregation_method = aggregation_method
        if aggregation_method == "attention":
            # same module tree as buyer_tower.py:32-36 so reference checkpoints load unchanged
            self.attention = nn.Sequential(
                nn.Linear(embedding_dim, attention_hidden_dim),
                nn.ReLU(),
                nn.Linear(attention_hidden_dim, 1),
            )
        elif aggregation_method == "weighted_avg":
            pass
        else:
            raise ValueError(f"Unknown aggregation method: {aggregation_method}")

    # -- helpers ---------------------------------------------------------------------------
    @staticmethod
    def _prep(item_embeddings: torch.Tensor, weights: torch.Tensor):
        if item_embeddings.dim() != 3 or weights.dim() != 2:
            raise ValueError("expected item_embeddings [B,S,D] and weights [B,S]")
        if item_embeddings.shape[:2] != weights.shape:
            raise ValueError(f"shape mismatch: {tuple(item_embeddings.shape)} vs {tuple(weights.shape)}")
        x = ops._f32c(item_embeddings, "item_embeddings")
        w = ops._f32c(weights.to(x.device) if weights.device != x.device else weights, "weights")
        return x, w

    def _mlp_params(self, device):
        l1, l2 = self.attention[0], self.attention[2]
        return (ops._f32c(l1.weight.detach().to(device), "attention.0.weight"),
                ops._f32c(l1.bias.detach().to(device), "attention.0.bias"),
                ops._f32c(l2.weight.detach().to(device).reshape(-1), "attention.2.weight"),
                ops._f32c(l2.bias.detach().to(device), "attention.2.bias"))

    # -- reference API ---------------------------------------------------------------------
    def weighted_average(self, item_embeddings: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
        """buyer_tower.py:43-68."""
        x, w = self._prep(item_embeddings, weights)
        return ops.pool_weighted(x, w).to(item_embeddings.dtype)

    def attention_aggregation(self, item_embeddings: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
        """buyer_tower.py:70-101."""
        x, w = self._prep(item_embeddings, weights)
        B, S, D = x.shape
        W1, b1, W2, b2 = self._mlp_params(x.device)
        logits = ops.attention_logits(x.view(B * S, D), W1, b1, W2, b2).view(B, S)
        return ops.pool_attention(x, logits, w).to(item_embeddings.dtype)

    def forward(self, item_embeddings: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
        """buyer_tower.py:103-122."""
        if self.aggregation_method == "weighted_avg":
            return self.weighted_average(item_embeddings, weights)
        elif self.aggregation_method == "attention":
            return self.attention_aggregation(item_embeddings, weights)
        else:
            raise ValueError(f"Unknown aggregation method: {self.aggregation_method}")

    def encode_from_sequence(self, item_embeddings: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
        """buyer_tower.py:124-144 — adds the batch dimension, returns [1, D]."""
        if item_embeddings.dim() == 2:
            item_embeddings = item_embeddings.unsqueeze(0)
        if weights.dim() == 1:
            weights = weights.unsqueeze(0)
        return self.forward(item_embeddings, weights)

    # -- device-resident item table (gather) path ------------------------------------------
    def precompute_item_logits(self, table: torch.Tensor) -> Optional[torch.Tensor]:
        """Attention logit of every item row ([N]); None in weighted_avg mode.  The logit depends on
        the item row only, so serving never runs the MLP again."""
        if self.aggregation_method != "attention":
            return None
        t = ops._f32c(table, "table")
        W1, b1, W2, b2 = self._mlp_params(t.device)
        return ops.attention_logits(t, W1, b1, W2, b2)

    def zero_row_logit(self) -> float:
        """Logit of an all-zero (padding) row: W2 . relu(b1) + b2."""
        l1, l2 = self.attention[0], self.attention[2]
        with torch.no_grad():
            return float((torch.relu(l1.bias.float()) * l2.weight.float().reshape(-1)).sum() + l2.bias.float()[0])

    def forward_gather(self, table: torch.Tensor, indices: torch.Tensor, weights: torch.Tensor,
                       item_logits: Optional[torch.Tensor] = None) -> torch.Tensor:
        """forward(table[indices], weights) without materialising the [B,S,D] gather.
        indices i64 [B,S]; out-of-range index = all-zero row (zero-padded history)."""
        t = ops._f32c(table, "table")
        idx = indices.to(device=t.device, dtype=torch.int64).contiguous()
        w = ops._f32c(weights.to(t.device), "weights")
        if idx.dim() != 2 or idx.shape != w.shape:
            raise ValueError("expected indices [B,S] and weights [B,S]")
        if self.aggregation_method == "weighted_avg":
            return ops.pool_weighted_gather(t, idx, w)
        elif self.aggregation_method == "attention":
            if item_logits is None:
                item_logits = self.precompute_item_logits(t)
            return ops.pool_attention_gather(t, ops._f32c(item_logits, "item_logits"), self.zero_row_logit(), idx, w)
        else:
            raise ValueError(f"Unknown aggregation method: {self.aggregation_method}")
