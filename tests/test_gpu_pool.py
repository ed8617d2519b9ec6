"""GPU: buyer-tower pooling kernels through the C-ABI vs the golden vectors of the real reference,
the CPU oracle on seeded inputs, and size-independent properties at the BASELINE C2 size."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN, golden, rel_err
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


@pytest.mark.parametrize("method", ["weighted_avg", "attention"])
def test_backward_matches_reference_formulation(method):
    """Training callers (two_tower.py:212): forward through the fused kernels, gradients equal to autograd of
    the reference arithmetic (x, and the attention MLP parameters)."""
    import two_tower_model_v2_b200 as pkg
    from two_tower_model_v2_b200.buyer_tower import _eager_pool
    torch.manual_seed(5)
    dev = torch.device("cuda:0")
    B, S, D = 6, 17, 384
    x = torch.randn(B, S, D, device=dev, requires_grad=True)
    w = torch.tensor([1.0, 5.0, 10.0], device=dev)[torch.randint(0, 3, (B, S), device=dev)]
    tower = pkg.BuyerTower(D, method).to(dev)
    out = tower(x, w)
    t = torch.randn_like(out)
    (out * t).sum().backward()
    gx = x.grad.clone()
    gp = [p.grad.clone() for p in tower.parameters()]
    x2 = x.detach().clone().requires_grad_(True)
    params = [p.detach().clone().requires_grad_(True) for p in tower.parameters()]
    ref = _eager_pool(x2, w, tuple(params) if params else None)
    (ref * t).sum().backward()
    assert torch.allclose(out, ref, atol=1e-6)
    assert torch.allclose(gx, x2.grad, atol=1e-6, rtol=1e-4)
    for a, b in zip(gp, params):
        assert torch.allclose(a, b.grad, atol=1e-6, rtol=1e-4)
    with torch.no_grad():                                   # inference path unchanged: no graph, same numbers
        assert torch.equal(tower(x, w), out.detach())
