"""CPU restatement of the reference BuyerTower arithmetic — TEST INFRASTRUCTURE ONLY.

Follows /root/reference/src/models/buyer_tower.py line by line in numpy (any float dtype, so the
same code gives the fp32 oracle and its fp64 cross-check), plus a torch-CPU port used as the timed
CPU baseline (the reference itself runs these ops in torch).

Pinned: tests/golden/buyer_tower_*.npz hold outputs of the *real* reference module (imported from
/root/reference in the authoring container by tests/golden/make_golden.py); tests/test_oracle.py
checks this restatement against them.
"""
from __future__ import annotations

import numpy as np


def l2_normalize(y: np.ndarray, eps: float = 1e-12) -> np.ndarray:
    """torch.nn.functional.normalize(y, p=2, dim=1): y / max(||y||_2, eps)  (buyer_tower.py:66,99)."""
    n = np.sqrt((y * y).sum(axis=1, keepdims=True))
    return y / np.maximum(n, np.asarray(eps, dtype=y.dtype))


def weighted_average(x: np.ndarray, w: np.ndarray) -> np.ndarray:
    """buyer_tower.py:43-68.  x [B,S,D], w [B,S] -> [B,D]."""
    w = w[..., None]                                              # :58  unsqueeze(-1)
    wsum = w.sum(axis=1, keepdims=True) + np.asarray(1e-8, dtype=x.dtype)   # :59
    nw = w / wsum                                                 # :60
    y = (x * nw).sum(axis=1)                                      # :63
    return l2_normalize(y)                                        # :66


def attention_logits(x: np.ndarray, W1, b1, W2, b2) -> np.ndarray:
    """buyer_tower.py:32-36,85-86.  Sequential(Linear(D,H), ReLU, Linear(H,1)) -> [B,S]."""
    h = np.maximum(x @ W1.T + b1, 0)
    return h @ W2.reshape(-1) + b2.reshape(())


def attention_aggregation(x: np.ndarray, w: np.ndarray, W1, b1, W2, b2) -> np.ndarray:
    """buyer_tower.py:70-101.  No masking: zero-weight events keep softmax mass e^0."""
    s = attention_logits(x, W1, b1, W2, b2)                       # :85-86
    c = s * w                                                     # :89
    c = c - c.max(axis=1, keepdims=True)
    e = np.exp(c)
    a = e / e.sum(axis=1, keepdims=True)                          # :92 softmax(dim=1)
    y = (x * a[..., None]).sum(axis=1)                            # :96
    return l2_normalize(y)                                        # :99


def forward(x, w, method: str, params=None) -> np.ndarray:
    """buyer_tower.py:103-122."""
    if method == "weighted_avg":
        return weighted_average(x, w)
    if method == "attention":
        return attention_aggregation(x, w, *params)
    raise ValueError(f"Unknown aggregation method: {method}")


def gather_rows(table: np.ndarray, idx: np.ndarray) -> np.ndarray:
    """History rows from an item table; out-of-range index = zero row (trainer.py:144-151 padding)."""
    ok = (idx >= 0) & (idx < table.shape[0])
    x = table[np.where(ok, idx, 0)]
    x[~ok] = 0
    return x


# ---- torch-CPU port: the timed CPU baseline (same ops the reference issues) -------------------
def torch_forward(x, w, method: str, params=None):
    import torch
    import torch.nn.functional as F
    if method == "weighted_avg":
        ww = w.unsqueeze(-1)
        nw = ww / (ww.sum(dim=1, keepdim=True) + 1e-8)
        return F.normalize((x * nw).sum(dim=1), p=2, dim=1)
    W1, b1, W2, b2 = params
    s = (torch.relu(x @ W1.t() + b1) @ W2.reshape(-1, 1) + b2).squeeze(-1)
    a = F.softmax(s * w, dim=1).unsqueeze(-1)
    return F.normalize((x * a).sum(dim=1), p=2, dim=1)
