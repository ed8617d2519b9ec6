"""Drop-in `InfoNCELoss` (reference: src/training/losses.py:8-79), the loss of the training callers of BuyerTower
(src/training/trainer.py:216-236).  Same constructor, same `forward(buyer_embeddings, positive_embeddings,
negative_embeddings) -> scalar`, differentiable; forward and backward run the fp32 CUDA kernels behind
tt_infonce_forward / tt_infonce_backward, which never build the reference's [B, B, D] expansion or its logits matrix.
CUDA float32 tensors only (there is no CPU path); SURVEY.md section 8f-4."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


class _InfoNCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, buyer, pos, neg, temperature):
        buyer, pos, neg = (ops._f32c(t, n) for t, n in ((buyer, "buyer_embeddings"), (pos, "positive_embeddings"),
                                                        (neg, "negative_embeddings")))
        loss, _, lse = ops.infonce_forward(buyer, pos, neg, temperature)
        ctx.save_for_backward(buyer, pos, neg, lse)
        ctx.temperature = temperature
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        buyer, pos, neg, lse = ctx.saved_tensors
        d_buyer, d_pos, d_neg = ops.infonce_backward(buyer, pos, neg, lse, g.reshape(1).contiguous().float(), ctx.temperature)
        return d_buyer, d_pos, d_neg, None


class InfoNCELoss(nn.Module):
    """InfoNCE (contrastive) loss with in-batch negatives + sampled negatives (losses.py:8-79)."""

    def __init__(self, temperature: float = 0.07):
        super().__init__()
        self.temperature = temperature

    def forward(self, buyer_embeddings: torch.Tensor, positive_embeddings: torch.Tensor,
                negative_embeddings: torch.Tensor) -> torch.Tensor:
        if buyer_embeddings.dim() != 2 or positive_embeddings.shape != buyer_embeddings.shape:
            raise ValueError("buyer_embeddings and positive_embeddings must both be [batch_size, embedding_dim]")
        if negative_embeddings.dim() != 3 or negative_embeddings.shape[0] != buyer_embeddings.shape[0] \
                or negative_embeddings.shape[2] != buyer_embeddings.shape[1]:
            raise ValueError("negative_embeddings must be [batch_size, num_negatives, embedding_dim]")
        return _InfoNCE.apply(buyer_embeddings, positive_embeddings, negative_embeddings, float(self.temperature))
