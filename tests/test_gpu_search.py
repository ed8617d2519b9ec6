"""GPU: exact inner-product search through the C-ABI vs the CPU oracle (seeded), the golden vectors
of the reference wrapper, and size-independent properties at the BASELINE 1M x 384 size."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN, golden
from oracle import flat_ip_oracle as fo

pytestmark = pytest.mark.gpu
VDB_CASES = sorted(p.name for p in GOLDEN.glob("vector_db_*.npz"))


def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def build(x):
    import two_tower_model_v2_b200 as pkg
    idx = pkg.FlatIPIndex(x.shape[1])
    idx.add(x)
    return idx


def check(idx, x, q, k, expect_certified=True):
    qd = torch.from_numpy(q).to(dev())
    scores, ids, flags, nunc = idx.search_device(qd, k)
    n_bad = int(nunc.item())
    if n_bad:
        qsel = torch.nonzero(flags != 1).flatten().to(torch.int32)
        idx.search_exact_device(qd, k, scores, ids, qsel)
    xn, qn = fo.normalize_rows(x), fo.normalize_rows(q)
    rs, ri = fo.search(xn, qn, k)
    ok, msg = fo.compare_topk(scores.cpu().numpy(), ids.cpu().numpy(), rs, ri, xn, qn)
    assert ok, msg
    if expect_certified:
        assert n_bad == 0, f"{n_bad}/{q.shape[0]} queries were not certified by the tensor-core path"
    return n_bad


# ---- the scan kernel itself: raw bf16 tensor-core scores vs the same arithmetic in torch ----------
@pytest.mark.parametrize("N,D,nq", [(256, 64, 128), (1000, 64, 5), (2048, 384, 128), (5000, 384, 1),
                                      (3000, 384, 200), (1500, 100, 70), (4096, 768, 64), (2000, 768, 100),
                                      (777, 512, 33), (1000, 1024, 64),
                                      (3000, 768, 300), (2500, 640, 129), (1800, 1024, 257), (1500, 704, 200)])
def test_scan_scores_match_bf16_reference(N, D, nq):
    rng = np.random.default_rng(N + D + nq)
    x = rng.standard_normal((N, D)).astype(np.float32)
    q = rng.standard_normal((nq, D)).astype(np.float32)
    idx = build(x)
    got = idx.scan_scores_device(torch.from_numpy(q).to(dev()))
    xh = idx.xh[:, :D].float()
    qn = torch.from_numpy(fo.normalize_rows(q)).to(dev())
    ref = (qn.to(torch.bfloat16).float().double() @ xh.double().T).float()
    assert not torch.isnan(got).any(), "the scan did not report every row"
    assert (got - ref).abs().max().item() < 2e-5   # fp32 accumulation-order noise only
    # stored rows: fp32 normalised copy and its bf16 rounding
    xn = fo.normalize_rows(x)
    assert np.abs(idx.xn.cpu().numpy() - xn).max() < 1e-6
    assert torch.equal(idx.xh[:, :D], idx.xn.to(torch.bfloat16))
    assert idx.xh[:, D:].abs().sum().item() == 0


@pytest.mark.parametrize("N,D,nq,k", [
    (20000, 384, 64, 10),      # C1 catalog / k
    (20000, 384, 2000, 10),    # C1 full query count (16 query blocks)
    (100000, 384, 1, 100), (100000, 384, 7, 100), (100000, 384, 129, 100), (50000, 384, 300, 1),
    (300000, 128, 1, 10), (100000, 384, 2, 100),   # nq = 1: threshold selected inside the main scan; nq = 2: radix-select kernel
    (30000, 768, 50, 100),     # C4 width (M=64 path)
    (60000, 768, 300, 100),    # C4 width, CTA pairs with a half-resident query block
    (30000, 100, 33, 37),      # D not a multiple of 64
    (9000, 64, 16, 1000),      # API maximum k (server.py:46)
    (70000, 256, 40, 500),
])
def test_search_matches_oracle(N, D, nq, k):
    rng = np.random.default_rng(N + nq + k)
    x = (rng.standard_normal((N, D)) * rng.uniform(0.2, 2.0, (N, 1))).astype(np.float32)
    q = rng.standard_normal((nq, D)).astype(np.float32)
    check(build(x), x, q, k)


@pytest.mark.parametrize("N,k", [(1, 1), (7, 7), (100, 10), (300, 300), (2048, 100), (2049, 100), (5000, 1000)])
def test_small_catalogs(N, k):
    rng = np.random.default_rng(N)
    x = rng.standard_normal((N, 32)).astype(np.float32)
    q = rng.standard_normal((9, 32)).astype(np.float32)
    check(build(x), x, q, k)


def test_duplicates_and_ties():
    """Duplicate rows tie exactly; ids inside a tie may differ from the oracle but must be valid, and our
    own order is (score desc, id asc)."""
    rng = np.random.default_rng(3)
    base = rng.standard_normal((500, 384)).astype(np.float32)
    x = np.concatenate([base] * 40)                      # every row appears 40 times
    q = base[:16] + 0.01 * rng.standard_normal((16, 384)).astype(np.float32)
    idx = build(x)
    check(idx, x, q, 100, expect_certified=False)
    s, i, _ = idx.search_checked_device(torch.from_numpy(q).to(dev()), 100)
    s, i = s.cpu().numpy(), i.cpu().numpy()
    for r in range(16):
        for a in range(99):
            assert s[r, a] > s[r, a + 1] or (s[r, a] == s[r, a + 1] and i[r, a] < i[r, a + 1])
    assert np.array_equal(i[:, :40] % 500, np.repeat(np.arange(16)[:, None], 40, 1))


def test_clustered_catalog_falls_back_exactly():
    """A catalog ordered by cluster defeats the sampled threshold for some queries; the certificate must
    catch it and the exact path must repair it."""
    rng = np.random.default_rng(4)
    centers = rng.standard_normal((50, 384)).astype(np.float32)
    x = np.repeat(centers, 2000, axis=0) + 0.05 * rng.standard_normal((100000, 384)).astype(np.float32)
    q = centers[:32] + 0.05 * rng.standard_normal((32, 384)).astype(np.float32)
    check(build(x), x, q, 100, expect_certified=False)


def test_exact_path_alone():
    rng = np.random.default_rng(8)
    x = rng.standard_normal((40000, 384)).astype(np.float32)
    q = rng.standard_normal((19, 384)).astype(np.float32)
    idx = build(x)
    s, i = idx.search_exact_device(torch.from_numpy(q).to(dev()), 100)
    xn, qn = fo.normalize_rows(x), fo.normalize_rows(q)
    rs, ri = fo.search(xn, qn, 100)
    ok, msg = fo.compare_topk(s.cpu().numpy(), i.cpu().numpy(), rs, ri, xn, qn)
    assert ok, msg


@pytest.mark.parametrize("case", VDB_CASES)
def test_vector_database_matches_reference_wrapper(case, tmp_path):
    import two_tower_model_v2_b200 as pkg
    g = golden(case)
    N, D = g["x"].shape
    pids = [f"p{i:05d}" for i in range(N)]
    db = pkg.VectorDatabase(embedding_dim=D)
    db.build_index(g["x"], pids)
    assert db.index.ntotal == N and db.id_to_index["p00003"] == 3 and db.index_to_id[3] == "p00003"
    k = int(g["k"])
    batch = db.retrieve_batch(g["q"], k)
    ids = np.array([[int(p[1:]) for p, _ in row] for row in batch])
    sc = np.array([[s for _, s in row] for row in batch], np.float32)
    assert ids.shape == g["batch_ids"].shape              # k = min(k, ntotal) clamp (vector_db.py:159)
    ok, msg = fo.compare_topk(sc, ids, g["batch_scores"], g["batch_ids"], fo.normalize_rows(g["x"]), fo.normalize_rows(g["q"]))
    assert ok, msg
    single = db.retrieve(g["q"][0], k)                    # 1-D query
    assert [int(p[1:]) for p, _ in single] == ids[0].tolist() and all(isinstance(s, float) for _, s in single)
    multi = db.retrieve(g["q"][:3], k)                    # [3,D] query: only row 0 comes back (vector_db.py:164)
    assert [p for p, _ in multi] == [p for p, _ in single]
    # save / load round trip through the faiss flat-file layout
    p_idx, p_ids, p_map = tmp_path / "product_index.faiss", tmp_path / "product_ids.npy", tmp_path / "product_id_to_index.json"
    db.save_index(str(p_idx), str(p_ids), str(p_map))
    db2 = pkg.VectorDatabase(embedding_dim=D)
    db2.load_index(str(p_idx), str(p_ids), str(p_map))
    assert db2.product_ids == pids and db2.index.ntotal == N
    assert db2.retrieve_batch(g["q"], k) == batch
    db3 = pkg.VectorDatabase(embedding_dim=D)
    db3.load_index(str(p_idx))                            # ids default to product_{i} (vector_db.py:84-86)
    assert db3.product_ids[5] == "product_5"


def test_topk_merge_matches_numpy():
    from two_tower_model_v2_b200 import ops
    rng = np.random.default_rng(2)
    G, nq, K = 8, 37, 100
    s = -np.sort(-rng.standard_normal((G, nq, K)).astype(np.float32), axis=2)
    s[3, :, 50:] = s[2, :, 50:]                            # cross-shard ties
    i = rng.permutation(G * nq * K).reshape(G, nq, K).astype(np.int64)
    s[5, 0, 90:], i[5, 0, 90:] = -np.inf, -1               # a short shard list
    ms, mi = ops.topk_merge(torch.from_numpy(s).to(dev()), torch.from_numpy(i).to(dev()))
    flat_s = s.transpose(1, 0, 2).reshape(nq, G * K)
    flat_i = i.transpose(1, 0, 2).reshape(nq, G * K)
    for r in range(nq):
        order = np.lexsort((flat_i[r], -flat_s[r]))[:K]
        assert np.array_equal(ms[r].cpu().numpy(), flat_s[r][order])
        assert np.array_equal(mi[r].cpu().numpy(), flat_i[r][order])


def test_1m_catalog_properties():
    """BASELINE config C3 size (1M x 384, k=100): every catalog row queried against the index finds itself
    first with score ~1, lists are sorted, everything is certified, and a sampled subset of queries
    matches the fp32 oracle."""
    import two_tower_model_v2_b200 as pkg
    N, D, k = 1_000_000, 384, 100
    g = torch.Generator(device="cuda").manual_seed(1234)
    x = torch.randn((N, D), device=dev(), generator=g)
    idx = pkg.FlatIPIndex.adopt(x)
    rows = torch.arange(0, N, 3907, device=dev())[:256]
    q = idx.xn[rows] * 2.5                                  # un-normalised copies of stored rows
    s, i, flags, nunc = idx.search_device(q, k)
    assert int(nunc.item()) == 0
    assert torch.equal(i[:, 0], rows) and (s[:, 0] - 1).abs().max() < 1e-5
    assert (s[:, 1:] <= s[:, :-1]).all()
    for nq in (1, 128, 1024):
        qr = torch.randn((nq, D), device=dev(), generator=g)
        s, i, flags, nunc = idx.search_device(qr, k)
        assert int(nunc.item()) == 0 and (s[:, 1:] <= s[:, :-1]).all()
        sub = slice(0, min(nq, 4))
        xn = idx.xn.cpu().numpy()
        qn = fo.normalize_rows(qr[sub].cpu().numpy())
        rs, ri = fo.search(xn, qn, k, block=1 << 18)
        ok, msg = fo.compare_topk(s[sub].cpu().numpy(), i[sub].cpu().numpy(), rs, ri, xn, qn)
        assert ok, msg


# ---- sharded path on one device: G shards of one catalog, records stacked as an all-gather would -----
def _shard_records(x, q, k, G):
    import two_tower_model_v2_b200 as pkg
    from two_tower_model_v2_b200.sharded import record_layout, record_views
    N, nq = x.shape[0], q.shape[0]
    lay = record_layout(nq, k)
    gathered = torch.zeros((G, lay.nbytes), dtype=torch.uint8, device=dev())
    qd = torch.from_numpy(q).to(dev())
    shards = []
    for g in range(G):
        lo, hi = pkg.shard_bounds(N, G, g)
        idx = build(x[lo:hi])
        idx.id_offset = lo
        s, i, b, f = record_views(gathered[g], lay, nq, k)
        kl = min(k, hi - lo)
        if kl < k:
            s.fill_(float("-inf")); i.fill_(-1)
        idx.search_shard_into(qd, kl, s, i, b, f)
        shards.append(idx)
    return gathered, lay, shards


@pytest.mark.parametrize("N,D,nq,k,G", [(40000, 64, 33, 100, 4), (70000, 384, 130, 10, 8), (900, 32, 5, 300, 3)])
def test_shard_merge_matches_oracle_and_numpy_merge(N, D, nq, k, G):
    from test_sharded_gloo import _merge_oracle
    from two_tower_model_v2_b200 import ops
    rng = np.random.default_rng(11)
    x = rng.standard_normal((N, D)).astype(np.float32)
    q = rng.standard_normal((nq, D)).astype(np.float32)
    gathered, lay, _ = _shard_records(x, q, k, G)
    s, i, flags, nunc = ops.shard_merge(gathered, lay.off_scores, lay.off_ids, lay.off_bound, lay.off_flags, nq, k)
    ns, ni, nf, nn = _merge_oracle(gathered.cpu(), lay, nq, k)
    assert torch.equal(s.cpu(), ns) and torch.equal(i.cpu(), ni) and torch.equal(flags.cpu(), nf)
    assert int(nunc.item()) == nn == 0
    xn, qn = fo.normalize_rows(x), fo.normalize_rows(q)
    rs, ri = fo.search(xn, qn, k)
    ok, msg = fo.compare_topk(s.cpu().numpy(), i.cpu().numpy(), rs, ri, xn, qn)
    assert ok, msg


def test_shard_merge_global_certificate_rejects_and_recovers():
    """A shard whose bound exceeds the merged K-th score must flag the query (reason 8); overflow (1) and
    self-check (4) bits of any shard propagate; local 'fewer than K' (2) alone does not."""
    from two_tower_model_v2_b200 import ops
    from two_tower_model_v2_b200.sharded import record_views
    rng = np.random.default_rng(12)
    N, D, nq, k, G = 30000, 64, 8, 50, 4
    x = rng.standard_normal((N, D)).astype(np.float32)
    q = rng.standard_normal((nq, D)).astype(np.float32)
    gathered, lay, _ = _shard_records(x, q, k, G)
    sg, ig, bg, fg = record_views(gathered, lay, nq, k)
    merged = ops.shard_merge(gathered, lay.off_scores, lay.off_ids, lay.off_bound, lay.off_flags, nq, k)
    kth = merged[0][:, k - 1].clone()
    bg[1, 0] = kth[0] + 1e-3          # shard 1 cannot exclude an unseen row above the K-th score of query 0
    bg[2, 1] = kth[1]                 # equality still certifies (unseen rows score strictly below the bound)
    fg[3, 2] = -1                     # overflow on shard 3, query 2
    fg[0, 3] = -4                     # self-check failure on shard 0, query 3
    fg[0, 4] = -2                     # locally short list only: superseded by the global check
    fg[1, 5] = -8                     # local K-th check only: superseded too
    s, i, flags, nunc = ops.shard_merge(gathered, lay.off_scores, lay.off_ids, lay.off_bound, lay.off_flags, nq, k)
    assert flags.cpu().tolist() == [-8, 1, -1, -4, 1, 1, 1, 1]
    assert int(nunc.item()) == 3


@pytest.mark.parametrize("N,D,nq,k,G", [(3_000_000, 64, 200, 100, 4), (2_200_000, 384, 33, 10, 8)])
def test_shard_global_threshold_path(N, D, nq, k, G):
    """tt_flat_shard_sample / tt_flat_shard_search on G shards of one device: one threshold from the gathered
    sample, ~target/G candidates per shard, merged result identical to the single-index search."""
    import two_tower_model_v2_b200 as pkg
    from two_tower_model_v2_b200 import _native, ops
    from two_tower_model_v2_b200.sharded import record_layout, record_views
    g = torch.Generator(device="cuda").manual_seed(77)
    x = torch.randn((N, D), device=dev(), generator=g)
    q = torch.randn((nq, D), device=dev(), generator=g)
    full = pkg.FlatIPIndex.adopt(x.clone())
    fs, fi, fbad = full.search_checked_device(q, k)
    shards = []
    for r in range(G):
        lo, hi = pkg.shard_bounds(N, G, r)
        idx = pkg.FlatIPIndex.adopt(x[lo:hi].clone())
        idx.id_offset = lo
        shards.append(idx)
    n_min = min(s.ntotal for s in shards)
    assert shards[0].shard_plan_ok(N, nq, k, n_min)
    topr_g = torch.empty((G, nq, _native.TT_SHARD_TOPR), device=dev())
    for r, idx in enumerate(shards):
        idx.shard_sample(q, k, N, topr_g[r])
    assert (topr_g[:, :, 1:] <= topr_g[:, :, :-1]).all()           # descending lists, -inf padded
    lay = record_layout(nq, k)
    gathered = torch.zeros((G, lay.nbytes), dtype=torch.uint8, device=dev())
    for r, idx in enumerate(shards):
        s, i, b, f = record_views(gathered[r], lay, nq, k)
        idx.shard_search_into(nq, k, N, topr_g, s, i, b, f)
    s, i, flags, nunc = ops.shard_merge(gathered, lay.off_scores, lay.off_ids, lay.off_bound, lay.off_flags, nq, k)
    assert int(nunc.item()) == 0
    assert torch.equal(i, fi) and torch.equal(s, fs)
    # every shard used the same threshold: the bounds differ only through the local prune cutoff
    _, _, bg, _ = record_views(gathered, lay, nq, k)
    assert torch.isfinite(bg).all()


@pytest.mark.parametrize("method", ["weighted_avg", "attention"])
def test_retrieval_pipeline_matches_encode_then_search_oracle(method):
    """history ids -> fused gather+pool -> search on the device == oracle pooling of the catalog rows + oracle search."""
    import two_tower_model_v2_b200 as pkg
    from oracle import buyer_tower_oracle as bo
    rng = np.random.default_rng(21)
    N, D, B, S, k = 30000, 384, 16, 50, 10
    cat = rng.standard_normal((N, D)).astype(np.float32)
    ids = [f"p{i}" for i in range(N)]
    db = pkg.VectorDatabase(D)
    db.build_index(cat, ids)
    torch.manual_seed(3)
    tower = pkg.BuyerTower(D, method).to(dev())
    pipe = pkg.RetrievalPipeline(tower, db)
    events = ["view", "add_to_cart", "purchase"]
    rows = rng.integers(0, N, (B, S))
    ev = rng.integers(0, 3, (B, S))
    lens = rng.integers(1, S + 1, B)
    batch = [[{"product_id": ids[rows[b, s]], "event_type": events[ev[b, s]]} for s in range(lens[b])] for b in range(B)]
    got = pipe.retrieve_batch(batch, k)
    xn = fo.normalize_rows(cat)
    wts = np.array([1.0, 5.0, 10.0], np.float32)
    for b in range(B):
        x = xn[rows[b, :lens[b]]][None]
        w = wts[ev[b, :lens[b]]][None]
        if method == "weighted_avg":
            emb = bo.weighted_average(x, w)
        else:
            emb = bo.attention_aggregation(x, w, *[p.detach().cpu().numpy() for p in tower.attention.parameters()])
        rs, ri = fo.search(xn, fo.normalize_rows(emb), k)
        assert [p for p, _ in got[b]] == [ids[i] for i in ri[0]], f"buyer {b}"
        assert np.allclose([s for _, s in got[b]], rs[0], atol=1e-5)
    one = pipe.retrieve(batch[0], k)
    assert [p for p, _ in one] == [p for p, _ in got[0]]


def test_torch_ops_build_search_exact():
    """The search is also reachable as torch.ops.tt.* (CUDA only): build -> search -> exact agree with the oracle."""
    import two_tower_model_v2_b200  # noqa: F401  (registers the ops)
    from two_tower_model_v2_b200 import _native
    rng = np.random.default_rng(31)
    N, D, nq, k = 50000, 96, 21, 10
    x = rng.standard_normal((N, D)).astype(np.float32)
    q = rng.standard_normal((nq, D)).astype(np.float32)
    xd = torch.from_numpy(x).to(dev())
    xn = torch.empty_like(xd)
    xh = torch.empty((N, int(_native.load().tt_flat_pitch(D))), device=dev(), dtype=torch.bfloat16)
    stats = torch.zeros(4, device=dev())
    torch.ops.tt.flat_build(xd, xn, xh, stats, 0, True)
    s, i, flags, nunc = torch.ops.tt.flat_search(torch.from_numpy(q).to(dev()), xn, xh, stats, k, 0)
    assert int(nunc.item()) == 0 and bool((flags == 1).all())
    es, ei = torch.ops.tt.flat_search_exact(torch.from_numpy(q).to(dev()), xn, k, 0)
    assert torch.equal(i, ei) and (s - es).abs().max().item() <= 1e-6
    xnn, qn = fo.normalize_rows(x), fo.normalize_rows(q)
    rs, ri = fo.search(xnn, qn, k)
    ok, msg = fo.compare_topk(s.cpu().numpy(), i.cpu().numpy(), rs, ri, xnn, qn)
    assert ok, msg


def test_native_shard_save_load_search_identical(tmp_path):
    """A shard written with save_native and read back with load_native answers bit-identically (no re-normalisation)."""
    import two_tower_model_v2_b200 as pkg
    rng = np.random.default_rng(41)
    x = rng.standard_normal((30000, 100)).astype(np.float32)
    q = torch.from_numpy(rng.standard_normal((17, 100)).astype(np.float32)).to(dev())
    idx = build(x)
    idx.id_offset = 777
    p = str(tmp_path / "shard.ttb2")
    idx.save_native(p, n_total=100000)
    back = pkg.FlatIPIndex.load_native(p)
    assert back.ntotal == idx.ntotal and back.id_offset == 777
    assert torch.equal(back.xn, idx.xn) and torch.equal(back.xh.view(torch.int16), idx.xh.view(torch.int16))
    s1, i1, _ = idx.search_checked_device(q, 10)
    s2, i2, _ = back.search_checked_device(q, 10)
    assert torch.equal(s1, s2) and torch.equal(i1, i2) and int(i1.min()) >= 777


def test_wide_rows_route_to_exact_path():
    """D > 1024 does not fit the resident-query scan: tt_flat_search serves it through the fp32 exact path
    (faiss has no such limit, vector_db.py:160)."""
    rng = np.random.default_rng(51)
    x = rng.standard_normal((3000, 1536)).astype(np.float32)
    q = rng.standard_normal((5, 1536)).astype(np.float32)
    check(build(x), x, q, 10)


def test_float64_catalog_is_normalised_in_float64_like_the_reference():
    """vector_db.py:44-45,51 normalises in the input dtype and then casts to f32."""
    import two_tower_model_v2_b200 as pkg
    rng = np.random.default_rng(52)
    x = rng.standard_normal((5000, 64)) * 3.0
    db = pkg.VectorDatabase(64)
    db.build_index(x, [str(i) for i in range(5000)])
    ref = (x / (np.linalg.norm(x, axis=1, keepdims=True) + 1e-8)).astype(np.float32)
    assert np.array_equal(db.index.xn.cpu().numpy(), ref)


def test_appending_rows_keeps_earlier_rows_and_results():
    """FlatIPIndex.add grows its storage geometrically; rows already stored and their order are kept."""
    rng = np.random.default_rng(53)
    x = rng.standard_normal((9000, 96)).astype(np.float32)
    q = rng.standard_normal((7, 96)).astype(np.float32)
    import two_tower_model_v2_b200 as pkg
    idx = pkg.FlatIPIndex(96)
    for lo in range(0, 9000, 1000):
        idx.add(x[lo:lo + 1000])
    assert idx.ntotal == 9000 and idx._xn_store.shape[0] >= 9000
    check(idx, x, q, 10)


def test_async_search_from_worker_thread_and_micro_batcher():
    """The completion events of search_async / search_host_async are recorded on the INDEX's stream, whatever the
    calling thread's current device is: a MicroBatcher worker thread (fresh thread, default device) must get
    complete, certified results."""
    import threading
    import two_tower_model_v2_b200 as pkg
    rng = np.random.default_rng(54)
    x = rng.standard_normal((50000, 64)).astype(np.float32)
    q = rng.standard_normal((40, 64)).astype(np.float32)
    idx = build(x)
    xn, qn = fo.normalize_rows(x), fo.normalize_rows(q)
    rs, ri = fo.search(xn, qn, 10)

    def batch_fn(payloads, k):
        s, i, _ = idx.search_host_async(np.stack(payloads), k).result()
        return pkg.ArrayRows(i, s)
    out = [None] * 40
    with pkg.MicroBatcher(batch_fn, max_batch=16, max_wait_ms=1.0) as mb:
        def client(c):
            for j in range(c, 40, 8):
                out[j] = mb(q[j], 10)
        ths = [threading.Thread(target=client, args=(c,)) for c in range(8)]
        [t.start() for t in ths]
        [t.join() for t in ths]
    ids = np.stack([o.ids for o in out])
    sc = np.stack([o.scores for o in out])
    ok, msg = fo.compare_topk(sc, ids, rs, ri, xn, qn)
    assert ok, msg


def test_index_on_non_current_device():
    """FlatIPIndex(d, device=cuda:1) used while cuda:0 is current: events / copies follow the index."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import two_tower_model_v2_b200 as pkg
    rng = np.random.default_rng(55)
    x = rng.standard_normal((30000, 64)).astype(np.float32)
    q = rng.standard_normal((9, 64)).astype(np.float32)
    torch.cuda.set_device(0)
    idx = pkg.FlatIPIndex(64, device=torch.device("cuda:1"))
    idx.add(x)
    s, i = idx.search(q, 10)
    xn, qn = fo.normalize_rows(x), fo.normalize_rows(q)
    rs, ri = fo.search(xn, qn, 10)
    ok, msg = fo.compare_topk(s, i, rs, ri, xn, qn)
    assert ok, msg
