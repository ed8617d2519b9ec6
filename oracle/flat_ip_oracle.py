"""CPU restatement of the exact inner-product index — TEST INFRASTRUCTURE ONLY.

The search arithmetic of the reference lives in a third-party dependency that is NOT under
/root/reference: `faiss-cpu>=1.7.4` (requirements.txt:26; no lockfile, no exact pin), used at
src/inference/vector_db.py:48 (IndexFlatIP), :54 (add), :160/:197 (search), :77/:117 (read/write).
faiss is not installable here, so this restates its published contract for IndexFlatIP: for each
query the exact fp32 inner product with every stored fp32 row, the k largest, sorted descending,
int64 labels.  **Parity at the faiss boundary is therefore UNPINNED** (the reference has no test or
golden vector touching VectorDatabase or faiss); what IS pinned is the wrapper logic around it
(normalisation, k clamp, id mapping) against the unmodified reference `vector_db.py` executed over
`oracle/faiss_shim` (tests/golden/vector_db_*.npz).

Tie policy: faiss's order among equal scores is implementation-defined (heap vs reservoir); the
north-star tolerance allows ids to differ inside score ties of 1e-6.  This oracle orders ties by
ascending row id and `compare_topk` implements that tolerance.
"""
from __future__ import annotations

import numpy as np


def normalize_rows(x: np.ndarray) -> np.ndarray:
    """vector_db.py:44-45,51 (and :152-156 / :189-193 for queries): x/(||x||+1e-8) in the input dtype, then f32."""
    n = np.linalg.norm(x, axis=1, keepdims=True)
    return (x / (n + 1e-8)).astype(np.float32)


def search(xn: np.ndarray, qn: np.ndarray, k: int, block: int = 65536, dtype=np.float32):
    """IndexFlatIP.search on already-normalised operands: (scores [nq,k], ids [nq,k] int64),
    descending score, ties by ascending id.  `dtype=np.float64` gives the cross-check."""
    nq = qn.shape[0]
    n = xn.shape[0]
    k = min(k, n)
    q = qn.astype(dtype)
    best_s = np.full((nq, 0), 0, dtype=dtype)
    best_i = np.zeros((nq, 0), dtype=np.int64)
    for lo in range(0, n, block):
        hi = min(n, lo + block)
        s = q @ xn[lo:hi].astype(dtype).T
        ids = np.broadcast_to(np.arange(lo, hi, dtype=np.int64), s.shape)
        cs = np.concatenate([best_s, s], axis=1)
        ci = np.concatenate([best_i, ids], axis=1)
        if cs.shape[1] > k:
            part = np.argpartition(-cs, k - 1, axis=1)[:, :k]
            # argpartition may cut through a tie group: pull in every element tied with the k-th
            kth = np.take_along_axis(cs, part, 1).min(axis=1, keepdims=True)
            keep_rows = []
            for r in range(nq):
                sel = np.nonzero(cs[r] >= kth[r])[0]
                order = np.lexsort((ci[r, sel], -cs[r, sel]))[:k]
                keep_rows.append(sel[order])
            sel = np.stack(keep_rows)
        else:
            sel = np.stack([np.lexsort((ci[r], -cs[r])) for r in range(nq)])
        best_s = np.take_along_axis(cs, sel, 1)
        best_i = np.take_along_axis(ci, sel, 1)
    return best_s.astype(np.float32), best_i


def reference_retrieve_batch(x: np.ndarray, q: np.ndarray, k: int):
    """VectorDatabase.build_index + retrieve_batch as arrays (vector_db.py:25-61,171-209)."""
    return search(normalize_rows(x), normalize_rows(q), min(k, x.shape[0]))


def compare_topk(scores, ids, ref_scores, ref_ids, xn=None, qn=None, score_tol=1e-5, tie_tol=1e-6):
    """North-star acceptance: scores within `score_tol`; ids identical except where the differing ids
    sit inside a score tie (|score - boundary score| <= tie_tol, judged on fp64 scores when the
    operands are given).  Returns (ok, message)."""
    scores, ids, ref_scores, ref_ids = map(np.asarray, (scores, ids, ref_scores, ref_ids))
    if scores.shape != ref_scores.shape or ids.shape != ref_ids.shape:
        return False, f"shape mismatch {scores.shape} vs {ref_scores.shape}"
    ds = np.abs(scores.astype(np.float64) - ref_scores.astype(np.float64))
    if ds.size and ds.max() > score_tol:
        return False, f"score error {ds.max():.3e} > {score_tol}"
    if np.any(np.diff(scores.astype(np.float64), axis=1) > 0):
        return False, "scores not sorted descending"
    bad = np.nonzero((ids != ref_ids).any(axis=1))[0]
    for r in bad:
        a, b = set(ids[r].tolist()), set(ref_ids[r].tolist())
        only = list((a - b) | (b - a))
        diff_pos = np.nonzero(ids[r] != ref_ids[r])[0]
        if xn is not None and qn is not None:
            ex = lambda rows: xn[np.asarray(rows, dtype=np.int64)].astype(np.float64) @ qn[r].astype(np.float64)
            if only:
                so = ex(only)
                kth = ex(ref_ids[r][-1:])[0]
                if np.abs(so - kth).max() > tie_tol + 1e-7:
                    return False, f"query {r}: id sets differ outside a tie (max gap {np.abs(so - kth).max():.3e})"
            # same set, different order: the swapped positions must be ties
            sa, sb = ex(ids[r][diff_pos]), ex(ref_ids[r][diff_pos])
            if np.abs(sa - sb).max() > tie_tol + 1e-7:
                return False, f"query {r}: order differs outside a tie ({np.abs(sa - sb).max():.3e})"
        else:
            if np.abs(scores[r][diff_pos] - ref_scores[r][diff_pos]).max() > tie_tol + 2e-7:
                return False, f"query {r}: ids differ outside a tie"
    return True, f"ok ({len(bad)} queries differ only inside ties)"


# ---- timed CPU baseline (port): what faiss-cpu does for nq >= 20, in torch ---------------------
def torch_search(xn_t, qn_t, k: int, block: int = 16384):
    """Blocked fp32 sgemm + running top-k (faiss: BLAS sgemm blocks + heap/reservoir). torch CPU tensors."""
    import torch
    nq = qn_t.shape[0]
    best_s = torch.full((nq, k), float("-inf"))
    best_i = torch.full((nq, k), -1, dtype=torch.int64)
    n = xn_t.shape[0]
    for lo in range(0, n, block):
        hi = min(n, lo + block)
        s = qn_t @ xn_t[lo:hi].t()
        kk = min(k, hi - lo)
        ts, ti = torch.topk(s, kk, dim=1)
        cs = torch.cat([best_s, ts], 1)
        ci = torch.cat([best_i, ti + lo], 1)
        best_s, sel = torch.topk(cs, k, dim=1)
        best_i = torch.gather(ci, 1, sel)
    return best_s, best_i
