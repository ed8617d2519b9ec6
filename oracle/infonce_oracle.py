"""CPU restatement of the reference InfoNCE loss — TEST INFRASTRUCTURE ONLY.

Follows /root/reference/src/training/losses.py:36-79 line by line in numpy (any float dtype: the same code gives the
fp32 oracle and its fp64 cross-check), plus the analytic gradients of that arithmetic (softmax - one-hot, chain rule
through the three dot-product blocks) and a differentiable torch restatement for autograd cross-checks.

Pinned: tests/golden/infonce_*.npz hold the loss AND the gradients of the *real* reference module
(`InfoNCELoss` imported from /root/reference and differentiated by torch autograd in the authoring container,
tests/golden/make_golden.py); tests/test_oracle.py checks this restatement against them.
"""
from __future__ import annotations

import numpy as np


def logits(b: np.ndarray, p: np.ndarray, n: np.ndarray, temperature: float) -> np.ndarray:
    """[B, 1 + M + B] logits of losses.py:39-71 (the masked diagonal of the in-batch block is -inf)."""
    B = b.shape[0]
    t = np.asarray(temperature, dtype=b.dtype)
    positive_sim = (b * p).sum(axis=1) / t                                  # :39-40
    negative_sim = np.einsum("bd,bmd->bm", b, n) / t                        # :45-48
    in_batch_sim = (b @ p.T) / t                                            # :54-60  (bmm against the expanded positives)
    in_batch_sim = np.where(np.eye(B, dtype=bool), -np.inf, in_batch_sim)   # :63-64
    all_negatives = np.concatenate([negative_sim, in_batch_sim], axis=1)    # :67
    return np.concatenate([positive_sim[:, None], all_negatives], axis=1)   # :70


def loss(b: np.ndarray, p: np.ndarray, n: np.ndarray, temperature: float = 0.07):
    """-> (loss scalar, row_loss [B], lse [B]): F.cross_entropy(logits, 0) with mean reduction (:73-77)."""
    lg = logits(b, p, n, temperature)
    m = lg.max(axis=1, keepdims=True)
    lse = (m + np.log(np.exp(lg - m).sum(axis=1, keepdims=True)))[:, 0]
    row = lse - lg[:, 0]
    return row.mean(), row, lse


def gradients(b: np.ndarray, p: np.ndarray, n: np.ndarray, temperature: float = 0.07, grad_loss: float = 1.0):
    """Analytic gradients of `loss` -> (d_buyer [B,D], d_pos [B,D], d_neg [B,M,D])."""
    B, M = b.shape[0], n.shape[1]
    lg = logits(b, p, n, temperature)
    m = lg.max(axis=1, keepdims=True)
    sm = np.exp(lg - m)
    sm /= sm.sum(axis=1, keepdims=True)
    sm[:, 0] -= 1.0
    w = sm * (grad_loss / B / temperature)                   # dL/d(dot product) of every (row, candidate)
    w_pos, w_neg, w_inb = w[:, 0], w[:, 1:1 + M], w[:, 1 + M:]      # w_inb[i, k] = weight of buyer i on p_k (0 on the diagonal)
    d_buyer = w_pos[:, None] * p + np.einsum("bm,bmd->bd", w_neg, n) + w_inb @ p
    d_pos = w_pos[:, None] * b + w_inb.T @ b
    d_neg = w_neg[:, :, None] * b[:, None, :]
    return d_buyer.astype(b.dtype), d_pos.astype(b.dtype), d_neg.astype(b.dtype)


def torch_loss(b, p, n, temperature: float = 0.07):
    """The same arithmetic in differentiable torch ops (for autograd cross-checks at any dtype)."""
    import torch
    import torch.nn.functional as F
    B = b.shape[0]
    pos = (b * p).sum(dim=1) / temperature
    neg = torch.bmm(b.unsqueeze(1), n.transpose(1, 2)).squeeze(1) / temperature
    inb = (b @ p.t()) / temperature
    inb = inb.masked_fill(torch.eye(B, device=b.device, dtype=torch.bool), float("-inf"))
    lg = torch.cat([pos.unsqueeze(1), neg, inb], dim=1)
    return F.cross_entropy(lg, torch.zeros(B, dtype=torch.long, device=b.device))
