"""GPU: buyer-tower pooling kernels through the C-ABI vs the golden vectors of the real reference,
the CPU oracle on seeded inputs, and size-independent properties at the BASELINE C2 size."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN, elem_rel_err, golden, rel_err
from oracle import buyer_tower_oracle as bo

pytestmark = pytest.mark.gpu
TOL = 1e-5   # north star: pooled embeddings within 1e-5 relative
BUYER_CASES = sorted(p.name for p in GOLDEN.glob("buyer_tower_*.npz"))


def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def make_tower(g, method):
    import two_tower_model_v2_b200 as pkg
    D, H = g["x"].shape[2], g["W1"].shape[0]
    m = pkg.BuyerTower(D, method, H).to(dev()).eval()
    if method == "attention":
        m.load_state_dict({"attention.0.weight": torch.from_numpy(g["W1"]), "attention.0.bias": torch.from_numpy(g["b1"]),
                           "attention.2.weight": torch.from_numpy(g["W2"]), "attention.2.bias": torch.from_numpy(g["b2"])})
    return m


@pytest.mark.parametrize("case", BUYER_CASES)
@pytest.mark.parametrize("method", ["weighted_avg", "attention"])
def test_pool_matches_reference_golden(case, method):
    g = golden(case)
    m = make_tower(g, method)
    x, w = torch.from_numpy(g["x"]).to(dev()), torch.from_numpy(g["w"]).to(dev())
    with torch.no_grad():
        out = m(x, w)
    assert out.shape == g[method].shape and out.dtype == torch.float32 and out.is_cuda
    assert rel_err(out.cpu().numpy(), g[method]) < TOL
    assert elem_rel_err(out.cpu().numpy(), g[method]) < TOL       # element-wise, floor = row rms
    if method == "attention":
        seq = m.encode_from_sequence(x[0], w[0])           # [S,D],[S] -> [1,D]
        assert tuple(seq.shape) == (1, g["x"].shape[2])
        assert rel_err(seq.detach().cpu().numpy(), g["attention_seq0"]) < TOL


def test_zero_weights_give_exact_zero():
    g = golden("buyer_tower_zero_weight_row.npz")
    m = make_tower(g, "weighted_avg")
    out = m(torch.from_numpy(g["x"]).to(dev()), torch.from_numpy(g["w"]).to(dev())).cpu().numpy()
    assert np.all(out[0] == 0)


@pytest.mark.parametrize("B,S,D,H", [(1, 100, 384, 128), (300, 17, 384, 128), (1024, 50, 384, 128),
                                       (33, 50, 768, 128), (65, 9, 100, 24), (40, 3, 30, 7), (5, 130, 256, 200)])
def test_pool_matches_oracle_seeded(B, S, D, H):
    import two_tower_model_v2_b200 as pkg
    rng = np.random.default_rng(B * 1000 + S)
    x = rng.standard_normal((B, S, D)).astype(np.float32)
    w = np.array([1.0, 5.0, 10.0], np.float32)[rng.choice(3, (B, S), p=[0.75, 0.18, 0.07])]
    xt, wt = torch.from_numpy(x).to(dev()), torch.from_numpy(w).to(dev())
    out = pkg.BuyerTower(D, "weighted_avg")(xt, wt).cpu().numpy()
    assert rel_err(out, bo.weighted_average(x, w)) < TOL
    torch.manual_seed(B + S)
    m = pkg.BuyerTower(D, "attention", H).to(dev())
    params = [p.detach().cpu().numpy() for p in m.attention.parameters()]
    out = m(xt, wt).detach().cpu().numpy()
    ref = bo.attention_aggregation(x.astype(np.float64), w.astype(np.float64), *[p.astype(np.float64) for p in params])
    assert rel_err(out, ref) < TOL


@pytest.mark.parametrize("B,S,D,H,scale", [
    (4096, 50, 384, 128, 1.0),      # C2
    (4096, 1, 384, 128, 1.0),       # S = 1: 64 buyers per MMA tile
    (600, 7, 384, 128, 1.0), (90, 100, 384, 128, 1.0), (20, 300, 384, 128, 1.0), (3, 2000, 384, 128, 1.0),
    (300, 17, 64, 24, 1.0), (500, 33, 128, 100, 1.0), (257, 50, 256, 128, 1.0), (149, 50, 320, 77, 1.0),
    (1000, 50, 384, 128, 20.0),     # x20 parameters: logits of order 10, softmax close to one-hot
    (1000, 50, 384, 128, 1e-3),     # tiny activations: fp16 `lo` pieces go subnormal, absolute error stays ~1e-8
])
def test_fused_attention_matches_fp64_oracle(B, S, D, H, scale):
    """tt_pool_attention_fused (one kernel: fp16 two-piece tensor-core MLP + softmax pooling + L2 norm) on shapes the
    fused path takes (B*S >= 4096, D % 64 == 0, D <= 384, H <= 128) vs the fp64 restatement of buyer_tower.py:70-101."""
    import two_tower_model_v2_b200 as pkg
    rng = np.random.default_rng(B + S + D)
    x = (rng.standard_normal((B, S, D)) * (scale if scale < 1 else 1.0)).astype(np.float32)
    w = np.array([1.0, 5.0, 10.0], np.float32)[rng.choice(3, (B, S), p=[0.75, 0.18, 0.07])]
    if S > 4:
        w[0, S // 2:] = 0           # zero-weight events keep their softmax mass (no masking in the reference)
        x[1, S // 2:] = 0           # zero-padded history
    torch.manual_seed(B + S)
    m = pkg.BuyerTower(D, "attention", H).to(dev())
    if scale > 1:
        with torch.no_grad():
            for prm in m.attention.parameters():
                prm.mul_(scale)
    params = [p.detach().cpu().numpy().astype(np.float64) for p in m.attention.parameters()]
    with torch.no_grad():
        out = m(torch.from_numpy(x).to(dev()), torch.from_numpy(w).to(dev())).cpu().numpy()
    ref = bo.attention_aggregation(x.astype(np.float64), w.astype(np.float64), *params)
    # Ill-conditioned cases (x20 parameters: logits of order 100 inside exp) are held to a small multiple of the error
    # the fp32 arithmetic of the reference itself makes there (fp32 restatement vs fp64), never less than TOL.
    ref32 = bo.attention_aggregation(x, w, *[p.astype(np.float32) for p in params])
    tol_n = max(TOL, 4 * rel_err(ref32, ref))
    tol_e = max(TOL, 4 * elem_rel_err(ref32, ref))
    assert rel_err(out, ref) < tol_n and elem_rel_err(out, ref) < tol_e, (rel_err(out, ref), elem_rel_err(out, ref), tol_n, tol_e)


def test_fused_attention_out_of_fp16_range_is_recomputed_in_fp32():
    """|x| * 16 > 65504 cannot be split into fp16 pieces: the kernel raises its device flag and the predicated fp32
    kernel recomputes the whole call (no host sync); results still match the oracle, including the affected buyers."""
    import two_tower_model_v2_b200 as pkg
    rng = np.random.default_rng(77)
    B, S, D, H = 200, 50, 384, 128
    x = rng.standard_normal((B, S, D)).astype(np.float32)
    x[17, 3, 100] = 9.0e4
    x[60, :, :] *= 5000.0
    w = np.array([1.0, 5.0, 10.0], np.float32)[rng.integers(0, 3, (B, S))]
    torch.manual_seed(5)
    m = pkg.BuyerTower(D, "attention", H).to(dev())
    with torch.no_grad():
        for prm in m.attention.parameters():
            prm.mul_(1e-3)             # keep the huge rows' logits finite so that the comparison is meaningful
    params = [p.detach().cpu().numpy().astype(np.float64) for p in m.attention.parameters()]
    with torch.no_grad():
        out = m(torch.from_numpy(x).to(dev()), torch.from_numpy(w).to(dev())).cpu().numpy()
    ref = bo.attention_aggregation(x.astype(np.float64), w.astype(np.float64), *params)
    assert np.isfinite(out).all() and rel_err(out, ref) < 1e-4      # fp32 arithmetic on 1e4-scale logits


@pytest.mark.parametrize("method", ["weighted_avg", "attention"])
def test_gather_variant_matches_dense(method):
    import two_tower_model_v2_b200 as pkg
    rng = np.random.default_rng(11)
    N, B, S, D = 5000, 257, 50, 384
    table = rng.standard_normal((N, D)).astype(np.float32)
    table /= np.linalg.norm(table, axis=1, keepdims=True)
    idx = rng.integers(0, N, (B, S))
    idx[0, 40:] = -1            # padded tail -> zero rows
    idx[1, :] = N + 5           # fully out of range
    w = np.array([1.0, 5.0, 10.0], np.float32)[rng.integers(0, 3, (B, S))]
    w[0, 40:] = 0
    torch.manual_seed(3)
    m = pkg.BuyerTower(D, method).to(dev())
    tt = torch.from_numpy(table).to(dev())
    out = m.forward_gather(tt, torch.from_numpy(idx).to(dev()), torch.from_numpy(w).to(dev())).cpu().numpy()
    x = bo.gather_rows(table, idx)
    params = [p.detach().cpu().numpy() for p in m.parameters()] if method == "attention" else None
    ref = bo.forward(x, w, method, params)
    assert np.abs(out - ref).max() < TOL
    dense = m(torch.from_numpy(x).to(dev()), torch.from_numpy(w).to(dev())).detach().cpu().numpy()
    assert np.abs(out - dense).max() < 2e-6


@pytest.mark.parametrize("method", ["weighted_avg", "attention"])
def test_c2_full_size_properties(method):
    """BASELINE config C2 (4096 buyers x 50 events x 384): unit norm, invariance to event order,
    weighted_avg invariance to weight scale; a sampled subset against the oracle."""
    import two_tower_model_v2_b200 as pkg
    B, S, D = 4096, 50, 384
    g = torch.Generator(device="cuda").manual_seed(99)
    x = torch.randn((B, S, D), device=dev(), generator=g)
    w = torch.tensor([1.0, 5.0, 10.0], device=dev())[torch.multinomial(torch.tensor([0.75, 0.18, 0.07], device=dev()), B * S, True, generator=g)].view(B, S)
    torch.manual_seed(0)
    m = pkg.BuyerTower(D, method).to(dev())
    out = m(x, w)
    assert torch.allclose(out.norm(dim=1), torch.ones(B, device=dev()), atol=1e-5)   # tests/test_buyer_tower.py:33-34
    perm = torch.randperm(S, device=dev())
    assert (m(x[:, perm].contiguous(), w[:, perm].contiguous()) - out).abs().max() < 5e-6
    if method == "weighted_avg":
        assert (m(x, w * 3.0) - out).abs().max() < 5e-6
    sel = torch.arange(0, B, 97, device=dev())
    params = [p.detach().cpu().numpy() for p in m.parameters()] if method == "attention" else None
    ref = bo.forward(x[sel].cpu().numpy(), w[sel].cpu().numpy(), method, params)
    assert rel_err(out[sel].detach().cpu().numpy(), ref) < TOL


def _reference_pool_fp64(x, w, params):
    """buyer_tower.py:58-66 / :85-99 restated HERE in differentiable fp64 torch (CPU) - independent of the package's
    own backward helpers - so that autograd of this is the gradient oracle."""
    import torch.nn.functional as F
    if params is None:
        nw = w.unsqueeze(-1) / (w.unsqueeze(-1).sum(dim=1, keepdim=True) + 1e-8)
        return F.normalize((x * nw).sum(dim=1), p=2, dim=1)
    W1, b1, W2, b2 = params
    s_ = (torch.relu(x @ W1.t() + b1) @ W2.t() + b2).squeeze(-1)
    a = torch.softmax(s_ * w, dim=1)
    return F.normalize((x * a.unsqueeze(-1)).sum(dim=1), p=2, dim=1)


@pytest.mark.parametrize("method", ["weighted_avg", "attention"])
@pytest.mark.parametrize("B,S,D", [(6, 17, 384), (300, 50, 384), (9, 33, 100), (5, 7, 30)])
def test_backward_matches_fp64_autograd_of_the_reference_formulation(method, B, S, D):
    """Training callers (two_tower.py:212): forward through the fused kernels, backward through tt_pool_backward (+ two
    GEMMs for the score MLP); gradients w.r.t. x, the event weights and the MLP parameters against fp64 autograd of the
    reference formulation restated in this file.  (D = 30 is not a multiple of 4: the eager fallback of the backward.)"""
    import two_tower_model_v2_b200 as pkg
    torch.manual_seed(5)
    dev = torch.device("cuda:0")
    x = torch.randn(B, S, D, device=dev, requires_grad=True)
    w = torch.tensor([1.0, 5.0, 10.0], device=dev)[torch.randint(0, 3, (B, S), device=dev)].requires_grad_(True)
    tower = pkg.BuyerTower(D, method).to(dev)
    out = tower(x, w)
    t = torch.randn_like(out)
    (out * t).sum().backward()
    x64 = x.detach().double().cpu().requires_grad_(True)
    w64 = w.detach().double().cpu().requires_grad_(True)
    p64 = [p.detach().double().cpu().requires_grad_(True) for p in tower.parameters()]
    ref = _reference_pool_fp64(x64, w64, p64 if p64 else None)
    (ref * t.double().cpu()).sum().backward()

    def close(a, b, what):
        a, b = a.detach().double().cpu(), b.detach()
        scale = float(b.abs().max()) + 1e-30
        err = float((a - b).abs().max()) / scale
        assert err < 2e-5, f"{what}: {err:.2e} of max |grad| {scale:.2e}"
    assert float((out.detach().double().cpu() - ref.detach()).abs().max()) < 2e-6
    close(x.grad, x64.grad, "dL/dx")
    close(w.grad, w64.grad, "dL/dw")
    for (n, p), q in zip(tower.named_parameters(), p64):
        close(p.grad, q.grad, f"dL/d{n}")
    with torch.no_grad():                                   # inference path unchanged: no graph, same numbers
        assert torch.equal(tower(x, w), out.detach())


@pytest.mark.parametrize("method", ["weighted_avg", "attention"])
@pytest.mark.parametrize("G,B,S,D", [(8, 257, 50, 384), (3, 1, 100, 384), (4, 40, 7, 64), (2, 1024, 50, 384)])
def test_sharded_partial_pooling_matches_unsharded(method, G, B, S, D):
    """Owner-computes pooling over a row-sharded item table (BASELINE config C5): per-shard partial records
    (tt_pool_partial_gather) stacked as an all-gather would, merged by tt_pool_partial_merge, against the oracle
    on the gathered rows and against the unsharded gather kernel."""
    import two_tower_model_v2_b200 as pkg
    from two_tower_model_v2_b200 import ops
    rng = np.random.default_rng(G * 100 + B)
    N = 6000
    table = rng.standard_normal((N, D)).astype(np.float32)
    table /= np.linalg.norm(table, axis=1, keepdims=True)
    idx = rng.integers(0, N, (B, S))
    idx[0, S // 2:] = -1                 # zero-padded tail
    if B > 1:
        idx[1, :] = N + 7                # a buyer with no valid row at all
    w = np.array([1.0, 5.0, 10.0], np.float32)[rng.integers(0, 3, (B, S))]
    w[0, S // 2:] = 0
    torch.manual_seed(7)
    m = pkg.BuyerTower(D, method).to(dev())
    tt = torch.from_numpy(table).to(dev())
    it, wt = torch.from_numpy(idx).to(dev()), torch.from_numpy(w).to(dev())
    logits = m.precompute_item_logits(tt)
    zero_logit = m.zero_row_logit() if method == "attention" else 0.0
    parts = []
    for g in range(G):
        lo, hi = pkg.shard_bounds(N, G, g)
        parts.append(ops.pool_partial_gather(tt[lo:hi].contiguous(), lo, N, g == 0,
                                             None if logits is None else logits[lo:hi].contiguous(), zero_logit, it, wt))
    out = ops.pool_partial_merge(torch.stack(parts), method == "attention").cpu().numpy()
    x = bo.gather_rows(table, idx)
    params = [p.detach().cpu().numpy() for p in m.parameters()] if method == "attention" else None
    ref = bo.forward(x.astype(np.float64), w.astype(np.float64), method, None if params is None else [p.astype(np.float64) for p in params])
    assert np.abs(out - ref).max() < TOL
    single = m.forward_gather(tt, it, wt, logits).cpu().numpy()
    assert np.abs(out - single).max() < 2e-6
