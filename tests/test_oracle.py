"""CPU: the oracle restatements against the golden vectors produced by the real reference."""
import numpy as np
import pytest

from conftest import GOLDEN, golden, rel_err
from oracle import buyer_tower_oracle as bo
from oracle import c_oracle
from oracle import flat_ip_oracle as fo

BUYER_CASES = sorted(p.name for p in GOLDEN.glob("buyer_tower_*.npz"))
VDB_CASES = sorted(p.name for p in GOLDEN.glob("vector_db_*.npz"))


def test_golden_present():
    assert len(BUYER_CASES) >= 8 and len(VDB_CASES) >= 3


@pytest.mark.parametrize("case", BUYER_CASES)
def test_buyer_oracle_matches_reference(case):
    g = golden(case)
    params = (g["W1"], g["b1"], g["W2"], g["b2"])
    wa = bo.weighted_average(g["x"], g["w"])
    at = bo.attention_aggregation(g["x"], g["w"], *params)
    # north-star tolerance for pooled embeddings: 1e-5 relative (fp32 noise measured at ~2e-7)
    assert rel_err(wa, g["weighted_avg"]) < 1e-5
    assert rel_err(at, g["attention"]) < 1e-5
    # fp64 restatement against the reference run in fp64: tight
    p64 = tuple(p.astype(np.float64) for p in params)
    assert rel_err(bo.weighted_average(g["x"].astype(np.float64), g["w"].astype(np.float64)), g["weighted_avg_f64"]) < 1e-12
    assert rel_err(bo.attention_aggregation(g["x"].astype(np.float64), g["w"].astype(np.float64), *p64), g["attention_f64"]) < 1e-10
    # encode_from_sequence adds the batch dim and returns [1, D] (buyer_tower.py:124-144)
    assert g["attention_seq0"].shape == (1, g["x"].shape[2])
    assert rel_err(at[:1], g["attention_seq0"]) < 1e-5


def test_buyer_oracle_zero_weights_give_zero_vector():
    g = golden("buyer_tower_zero_weight_row.npz")
    assert np.all(g["weighted_avg"][0] == 0)
    assert np.all(bo.weighted_average(g["x"], g["w"])[0] == 0)


def test_torch_port_matches_numpy_oracle():
    import torch
    g = golden("buyer_tower_b8_s50_d384.npz")
    t = lambda a: torch.from_numpy(a)
    params = tuple(t(g[k]) for k in ("W1", "b1", "W2", "b2"))
    assert rel_err(bo.torch_forward(t(g["x"]), t(g["w"]), "weighted_avg").numpy(), g["weighted_avg"]) < 1e-6
    assert rel_err(bo.torch_forward(t(g["x"]), t(g["w"]), "attention", params).numpy(), g["attention"]) < 1e-6


def test_gather_rows_zero_pads():
    table = np.arange(12, dtype=np.float32).reshape(4, 3)
    x = bo.gather_rows(table, np.array([[0, 3, -1, 4]]))
    assert np.all(x[0, 0] == table[0]) and np.all(x[0, 1] == table[3]) and np.all(x[0, 2:] == 0)


@pytest.mark.parametrize("case", VDB_CASES)
def test_search_oracle_matches_reference_wrapper(case):
    g = golden(case)
    s, i = fo.reference_retrieve_batch(g["x"], g["q"], int(g["k"]))
    assert i.shape == g["batch_ids"].shape           # includes the k = min(k, ntotal) clamp
    ok, msg = fo.compare_topk(s, i, g["batch_scores"], g["batch_ids"], fo.normalize_rows(g["x"]), fo.normalize_rows(g["q"]))
    assert ok, msg
    assert np.array_equal(g["single_ids"], g["batch_ids"][0])
    assert np.array_equal(g["multi_as_single_ids"], g["batch_ids"][0])   # only query 0 (vector_db.py:164)


def test_c_oracle_matches_numpy_oracle():
    rng = np.random.default_rng(5)
    x = fo.normalize_rows(rng.standard_normal((5000, 96)).astype(np.float32))
    q = fo.normalize_rows(rng.standard_normal((19, 96)).astype(np.float32))
    for k in (1, 10, 100):
        s1, i1 = fo.search(x, q, k)
        s2, i2 = c_oracle.search(x, q, k, nthreads=3)
        ok, msg = fo.compare_topk(s2, i2, s1, i1, x, q)
        assert ok, msg
    assert rel_err(c_oracle.normalize_rows(x * 3.0), x) < 1e-6


def test_fp64_crosscheck_bounds_fp32_noise():
    rng = np.random.default_rng(6)
    x = fo.normalize_rows(rng.standard_normal((4000, 384)).astype(np.float32))
    q = fo.normalize_rows(rng.standard_normal((8, 384)).astype(np.float32))
    s32, i32 = fo.search(x, q, 50)
    s64, i64 = fo.search(x, q, 50, dtype=np.float64)
    assert np.abs(s32 - s64).max() < 1e-6
    ok, msg = fo.compare_topk(s32, i32, s64, i64, x, q)
    assert ok, msg


def test_ties_are_ordered_by_id_and_tolerated():
    x = np.zeros((6, 4), np.float32)
    x[:, 0] = 1.0                      # six identical rows -> all scores tie
    q = np.array([[1.0, 0, 0, 0]], np.float32)
    s, i = fo.search(x, q, 3)
    assert i.tolist() == [[0, 1, 2]]
    ok, _ = fo.compare_topk(s, np.array([[5, 4, 3]]), s, i, x, q)   # different members of the tie: accepted
    assert ok
    x2 = x.copy(); x2[5, 0] = 0.5
    s2, i2 = fo.search(x2, q, 3)
    ok, _ = fo.compare_topk(np.array([[1, 1, 0.5]], np.float32), np.array([[0, 1, 5]]), s2, i2, x2, q)
    assert not ok                      # a non-tied intruder is rejected


@pytest.mark.parametrize("name", ["buyer_tower_b8_s50_d384.npz", "buyer_tower_ragged.npz", "buyer_tower_zero_weight_row.npz"])
def test_eager_backward_formulation_matches_reference_outputs(name):
    """The differentiable restatement used by the backward pass (buyer_tower._eager_pool) reproduces the real
    reference module's outputs (golden vectors), so its autograd gradients are the reference's gradients."""
    import torch
    from two_tower_model_v2_b200.buyer_tower import _eager_pool
    g = golden(name)
    x, w = torch.from_numpy(g["x"]), torch.from_numpy(g["w"])
    wa = _eager_pool(x, w).numpy()
    assert rel_err(wa, g["weighted_avg"]) < 1e-5
    mlp = tuple(torch.from_numpy(g[k]) for k in ("W1", "b1", "W2", "b2"))
    at = _eager_pool(x, w, mlp).numpy()
    assert rel_err(at, g["attention"]) < 1e-5


INFONCE_CASES = sorted(p.name for p in GOLDEN.glob("infonce_*.npz"))


@pytest.mark.parametrize("case", INFONCE_CASES)
def test_infonce_oracle_matches_reference(case):
    """oracle/infonce_oracle.py (losses.py:36-79) against the loss and the autograd gradients of the real InfoNCELoss."""
    import torch
    from oracle import infonce_oracle as io
    g = golden(case)
    b, p, n, t = g["buyer"], g["pos"], g["neg"], float(g["temperature"])
    loss, row, lse = io.loss(b, p, n, t)
    assert abs(loss - g["loss"]) <= 1e-5 * max(1.0, abs(g["loss"]))
    l64, _, _ = io.loss(b.astype(np.float64), p.astype(np.float64), n.astype(np.float64), t)
    assert abs(l64 - g["loss_f64"]) <= 1e-10 * max(1.0, abs(g["loss_f64"]))
    assert row.shape == (b.shape[0],) and np.allclose(row.mean(), loss)
    db, dp, dn = io.gradients(b.astype(np.float64), p.astype(np.float64), n.astype(np.float64), t)
    for mine, ref in ((db, g["d_buyer"]), (dp, g["d_pos"]), (dn, g["d_neg"])):
        assert mine.shape == ref.shape
        if ref.size:
            assert np.abs(mine - ref).max() <= 1e-5 * max(np.abs(ref).max(), 1e-30)
    tl = io.torch_loss(torch.from_numpy(b), torch.from_numpy(p), torch.from_numpy(n), t)
    assert abs(tl.item() - g["loss"]) <= 1e-5 * max(1.0, abs(g["loss"]))


def test_infonce_golden_present():
    assert len(INFONCE_CASES) >= 5
