"""Buyer tower: drop-in for the reference `src/models/buyer_tower.py:9-144`.

Same constructor, attributes (`embedding_dim`, `aggregation_method`, `attention`), state-dict
keys (`attention.0.weight`, `attention.0.bias`, `attention.2.weight`, `attention.2.bias`),
method names, argument meaning and exceptions.  The arithmetic runs in hand-written sm_100a CUDA
kernels (csrc/pool.cu, csrc/attn_logits.cu) through the C-ABI; there is no CPU path: CPU tensors
raise RuntimeError.

Training callers work too: under autograd the forward still runs the fused kernels and the backward runs
tt_pool_backward plus two plain GEMMs for the score MLP (`_FusedPool`).

Beyond the reference surface: `precompute_item_logits` / `forward_gather` pool
straight out of a device-resident item-embedding table given history row indices, so a request
never re-encodes history items (reference: src/inference/encoder.py:276-303).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import ops


def _eager_pool(x: torch.Tensor, w: torch.Tensor, mlp=None) -> torch.Tensor:
    """The reference arithmetic in differentiable torch ops (buyer_tower.py:58-66 / :85-99).  Used ONLY inside the
    backward pass of shapes the fused backward kernel does not take (D % 4 != 0 or D > 1024); the forward always
    runs the fused CUDA kernels."""
    if mlp is None:
        nw = w.unsqueeze(-1) / (w.unsqueeze(-1).sum(dim=1, keepdim=True) + 1e-8)
        y = (x * nw).sum(dim=1)
    else:
        W1, b1, W2, b2 = mlp
        logits = torch.relu(x @ W1.t() + b1) @ W2.reshape(-1, 1) + b2
        a = torch.softmax(logits.squeeze(-1) * w, dim=1)
        y = (x * a.unsqueeze(-1)).sum(dim=1)
    return torch.nn.functional.normalize(y, p=2, dim=1)


class _FusedPool(torch.autograd.Function):
    """Forward: fused sm_100a kernels (no [B,S,D] temporary, no graph).  Backward (training callers of the
    reference: src/models/two_tower.py:212, src/training/trainer.py:216-236): tt_pool_backward - one fused kernel
    for the normalisation / softmax / weighted-sum part (dx, dw, dlogit) - plus, in attention mode, the score MLP's
    own backward as two plain GEMMs (autograd over the recomputed Linear-ReLU-Linear, a library GEMM)."""

    @staticmethod
    def forward(ctx, x, w, *mlp):
        ctx.save_for_backward(x, w, *mlp)
        if mlp:
            return ops.pool_attention_fused(x, w, mlp[0].contiguous(), mlp[1].contiguous(),
                                            mlp[2].reshape(-1).contiguous(), mlp[3].contiguous())
        return ops.pool_weighted(x, w)

    @staticmethod
    def backward(ctx, g):
        x, w, *mlp = ctx.saved_tensors
        needs = ctx.needs_input_grad
        B, S, D = x.shape
        if D % 4 != 0 or D > 1024:
            with torch.enable_grad():
                ins = [t.detach().requires_grad_(needs[i]) for i, t in enumerate((x, w, *mlp))]
                out = _eager_pool(ins[0], ins[1], tuple(ins[2:]) if mlp else None)
                wanted = [t for t in ins if t.requires_grad]
                grads = iter(torch.autograd.grad(out, wanted, g.contiguous()))
            return tuple(next(grads) if t.requires_grad else None for t in ins)
        g = g.contiguous().float()
        if not mlp:
            dx, dw, _ = ops.pool_backward(x, w, None, g, need_dx=needs[0])
            return (dx if needs[0] else None, dw if needs[1] else None)
        W1, b1, W2, b2 = mlp
        logits = ops.attention_logits(x.view(B * S, D), W1.contiguous(), b1.contiguous(), W2.reshape(-1).contiguous(),
                                      b2.contiguous()).view(B, S)
        dx, dw, dlogit = ops.pool_backward(x, w, logits, g, need_dx=needs[0])
        with torch.enable_grad():          # d logits / d (x, W1, b1, W2, b2): plain GEMMs
            ins = [x.detach().requires_grad_(needs[0])] + [t.detach().requires_grad_(needs[2 + i]) for i, t in enumerate(mlp)]
            lg = (torch.relu(ins[0] @ ins[1].t() + ins[2]) @ ins[3].reshape(-1, 1) + ins[4]).squeeze(-1)
            wanted = [t for t in ins if t.requires_grad]
            grads = iter(torch.autograd.grad(lg, wanted, dlogit)) if wanted else iter(())
        mg = [next(grads) if t.requires_grad else None for t in ins]
        gx = (dx + mg[0]) if needs[0] else None
        return (gx, dw if needs[1] else None, *mg[1:])


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
        if torch.is_grad_enabled() and (x.requires_grad or w.requires_grad):
            return _FusedPool.apply(x, w).to(item_embeddings.dtype)
        return ops.pool_weighted(x, w).to(item_embeddings.dtype)

    def attention_aggregation(self, item_embeddings: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
        """buyer_tower.py:70-101."""
        x, w = self._prep(item_embeddings, weights)
        B, S, D = x.shape
        l1, l2 = self.attention[0], self.attention[2]
        if torch.is_grad_enabled() and (x.requires_grad or w.requires_grad or l1.weight.requires_grad):
            # training: keep the parameters in the graph (the reference trains this MLP, trainer.py:216-236)
            return _FusedPool.apply(x, w, l1.weight.float(), l1.bias.float(), l2.weight.float(), l2.bias.float()) \
                .to(item_embeddings.dtype)
        W1, b1, W2, b2 = self._mlp_params(x.device)
        return ops.pool_attention_fused(x, w, W1, b1, W2, b2).to(item_embeddings.dtype)

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
