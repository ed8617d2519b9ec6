"""Soundness of the exactness certificates, checked on a numpy model of finalize + shard merge.

The CUDA kernels implement this logic; here the LOGIC itself is attacked with adversarial inputs: exact scores f,
"tensor-core" scores b with |b - f| <= eps chosen by hypothesis (including the extremes), arbitrary thresholds and
prune margins.  Claim under test: whenever the certificate passes, the emitted list IS the true top-K of f in (score desc, id asc)
order, exact ties included."""
import numpy as np
from hypothesis import given, settings, strategies as st


def finalize_model(f, b, eps, thr, K, margin):
    """One shard: returns (ids of the emitted list sorted by (f desc, id asc), certified?, bound)."""
    cand = np.flatnonzero(b >= thr)
    if len(cand) > K:
        bK = np.sort(b[cand])[::-1][K - 1]
        cutoff = bK - margin * eps
    else:
        cutoff = -np.inf
    surv = cand[b[cand] >= cutoff]
    order = surv[np.lexsort((surv, -f[surv]))]
    out = order[:K]
    c = max(thr, cutoff)
    bound = -np.inf if c == -np.inf else c + eps
    certified = len(surv) >= K and (bound == -np.inf or f[out[K - 1]] >= bound)
    return out, certified, bound


def true_topk(f, K):
    order = np.lexsort((np.arange(len(f)), -f))
    return order[:K]


@st.composite
def instance(draw):
    n = draw(st.integers(5, 120))
    K = draw(st.integers(1, min(n, 12)))
    eps = draw(st.sampled_from([0.0, 1e-3, 4e-3, 5e-2]))
    # scores on a coarse grid so that near-ties and exact ties at the boundary are common
    f = np.array(draw(st.lists(st.integers(-40, 40), min_size=n, max_size=n)), dtype=np.float64) * 0.0025
    sign = np.array(draw(st.lists(st.sampled_from([-1.0, -0.5, 0.0, 0.5, 1.0]), min_size=n, max_size=n)))
    b = f + sign * eps
    thr = draw(st.sampled_from([-np.inf, -0.05, 0.0, 0.02, 0.05, 0.08]))
    margin = draw(st.sampled_from([0.0, 0.5, 1.0, 1.5, 2.0, 3.0]))
    return f, b, eps, thr, K, margin


@settings(max_examples=1500, deadline=None)
@given(instance())
def test_single_shard_certificate_is_sound(inst):
    f, b, eps, thr, K, margin = inst
    out, ok, _ = finalize_model(f, b, eps, thr, K, margin)
    if ok:
        want = true_topk(f, K)
        # a row that was not rescored scores STRICTLY below the certified K-th score, so even exact ties at the
        # boundary are resolved among rescored rows only: ids and order match the (score desc, id asc) top-K
        assert np.array_equal(out, want)


@settings(max_examples=1000, deadline=None)
@given(instance(), st.integers(2, 5), st.data())
def test_sharded_global_certificate_is_sound(inst, G, data):
    """Shards finalize independently (own thresholds allowed), the merge takes the top-K of the union of their lists and
    certifies iff the merged K-th score clears every shard's bound of the rows it did not rescore."""
    f, b, eps, thr, K, margin = inst
    n = len(f)
    per = -(-n // G)
    merged, bounds, enough = [], [], True
    for g in range(G):
        lo, hi = min(n, g * per), min(n, (g + 1) * per)
        if hi <= lo:
            continue
        thr_g = data.draw(st.sampled_from([thr, -np.inf, 0.03]))
        k_local = min(K, hi - lo)
        out, _, bound = finalize_model(f[lo:hi], b[lo:hi], eps, thr_g, k_local, margin)
        merged.extend((out + lo).tolist())
        bounds.append(bound)
    merged = np.array(merged, dtype=np.int64)
    if len(merged) < K:
        return                                                   # flagged "fewer than K": re-run exactly
    order = merged[np.lexsort((merged, -f[merged]))][:K]
    fK = f[order[K - 1]]
    certified = all(fK >= bd for bd in bounds)
    if certified:
        assert np.array_equal(order, true_topk(f, K))


def test_certificate_rejects_when_threshold_hides_a_winner():
    """A row whose tensor-core score fell just under the threshold but whose exact score beats the K-th: must not certify."""
    f = np.array([0.50, 0.40, 0.30, 0.29, 0.10])
    b = f.copy()
    b[1] = 0.40 - 4e-3                  # under-estimated by eps
    out, ok, _ = finalize_model(f, b, 4e-3, 0.399, 2, 1.5)      # thr hides row 1; candidates = rows {0}
    assert not ok
    out, ok, _ = finalize_model(f, b, 4e-3, 0.25, 2, 1.5)       # thr low enough: rows 0..3 candidates
    assert ok and out.tolist() == [0, 1]
