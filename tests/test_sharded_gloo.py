"""CPU, world_size 2, gloo: the shard / all-gather / merge plumbing of the multi-GPU search, with the
local search and the merge replaced by the numpy oracle (the CUDA kernels are covered by -m gpu)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import flat_ip_oracle as fo


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _LocalOracleIndex:
    """Stands in for FlatIPIndex on a CPU rank: same search_shard_into / search_exact_into contract."""

    def __init__(self, xn_shard, id_offset):
        self.xn, self.id_offset = xn_shard, id_offset
        self.exact_calls = 0

    @property
    def ntotal(self):
        return self.xn.shape[0]

    def search_shard_into(self, q, k, scores, ids, bound, flags):
        s, i = fo.search(self.xn, fo.normalize_rows(q.numpy()), k)
        scores[:, :k] = torch.from_numpy(s)
        ids[:, :k] = torch.from_numpy(i + self.id_offset)
        bound.fill_(float("-inf"))
        flags.fill_(1)

    def search_exact_into(self, q, k, scores, ids, qsel):
        self.exact_calls += 1
        rows = qsel.long()
        s, i = fo.search(self.xn, fo.normalize_rows(q.numpy()[rows.numpy()]), k)
        scores[rows, :k] = torch.from_numpy(s)
        ids[rows, :k] = torch.from_numpy(i + self.id_offset)


class _GlobalThresholdOracleIndex(_LocalOracleIndex):
    """Adds the two-phase contract (shard_sample -> all-gather -> shard_search_into): the stand-in publishes
    its rank in the list and checks that it sees every rank's list before searching."""

    def __init__(self, xn_shard, id_offset, rank, world):
        super().__init__(xn_shard, id_offset)
        self.rank, self.world, self.q = rank, world, None

    def shard_plan_ok(self, n_total, nq, k, n_local_min):
        return n_local_min >= k

    def shard_sample(self, q, k, n_total, topr):
        self.q = q
        topr.fill_(float(self.rank))

    def shard_search_into(self, nq, k, n_total, topr_g, scores, ids, bound, flags):
        assert topr_g.shape[0] == self.world
        assert [float(topr_g[g, 0, 0]) for g in range(self.world)] == [float(g) for g in range(self.world)]
        self.search_shard_into(self.q, k, scores, ids, bound, flags)


def _merge_oracle(gathered, lay, nq, K):
    """numpy restatement of tt_shard_merge (merge + global certificate)."""
    from two_tower_model_v2_b200.sharded import record_views
    sg, ig, bg, fg = record_views(gathered, lay, nq, K)
    G = sg.shape[0]
    s = sg.permute(1, 0, 2).reshape(nq, G * K).numpy()
    i = ig.permute(1, 0, 2).reshape(nq, G * K).numpy()
    out_s, out_i = np.full((nq, K), -np.inf, np.float32), np.full((nq, K), -1, np.int64)
    flags = np.ones(nq, np.int32)
    for r in range(nq):
        valid = np.flatnonzero(i[r] >= 0)
        order = valid[np.lexsort((i[r][valid], -s[r][valid]))][:K]
        out_s[r, :len(order)], out_i[r, :len(order)] = s[r][order], i[r][order]
        why = 0
        for g in range(G):
            f = int(fg[g, r])
            if f <= 0:
                why |= (-f) & 5
        if len(valid) < K:
            why |= 2
        elif not (out_s[r, K - 1] >= float(bg[:, r].max())):
            why |= 8
        flags[r] = 1 if why == 0 else -why
    return torch.from_numpy(out_s), torch.from_numpy(out_i), torch.from_numpy(flags), int((flags != 1).sum())


class _FlagOnceMerge:
    """Merge that rejects query 2 the first time it is called: drives the exact re-run path."""

    def __init__(self):
        self.calls = 0

    def __call__(self, gathered, lay, nq, K):
        s, i, f, n = _merge_oracle(gathered, lay, nq, K)
        self.calls += 1
        if self.calls == 1:
            f = f.clone()
            f[2] = -8
            n += 1
        return s, i, f, n


def _worker(rank, world, port, N, D, k, out_dir, flag_once=False, two_phase=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import two_tower_model_v2_b200 as pkg
    rng = np.random.default_rng(0)                      # same catalog and queries on every rank
    xn = fo.normalize_rows(rng.standard_normal((N, D)).astype(np.float32))
    q = rng.standard_normal((6, D)).astype(np.float32)
    lo, hi = pkg.shard_bounds(N, world, rank)
    local = _GlobalThresholdOracleIndex(xn[lo:hi], lo, rank, world) if two_phase else _LocalOracleIndex(xn[lo:hi], lo)
    merge = _FlagOnceMerge() if flag_once else _merge_oracle
    sharded = pkg.ShardedFlatIPIndex(local, N, merge=merge)
    s, i, n_bad = sharded.search_device(torch.from_numpy(q), k)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), s=s.numpy(), i=i.numpy(), n_bad=n_bad, exact_calls=local.exact_calls)
    dist.destroy_process_group()


def test_sharded_search_world2(tmp_path):
    N, D, k, world = 1001, 24, 20, 2
    mp.spawn(_worker, args=(world, _free_port(), N, D, k, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(0)
    xn = fo.normalize_rows(rng.standard_normal((N, D)).astype(np.float32))
    q = rng.standard_normal((6, D)).astype(np.float32)
    rs, ri = fo.search(xn, fo.normalize_rows(q), k)
    for r in range(world):
        g = np.load(tmp_path / f"r{r}.npz")
        assert np.array_equal(g["i"], ri) and np.allclose(g["s"], rs, atol=1e-6)


def test_shard_smaller_than_k_is_padded(tmp_path):
    N, D, k, world = 30, 8, 20, 2                        # 15 rows per shard < k
    mp.spawn(_worker, args=(world, _free_port(), N, D, k, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(0)
    xn = fo.normalize_rows(rng.standard_normal((N, D)).astype(np.float32))
    q = rng.standard_normal((6, D)).astype(np.float32)
    rs, ri = fo.search(xn, fo.normalize_rows(q), k)
    g = np.load(tmp_path / "r0.npz")
    assert np.array_equal(g["i"], ri)


def test_uncertified_queries_are_rerun_exactly(tmp_path):
    N, D, k, world = 600, 16, 10, 2
    mp.spawn(_worker, args=(world, _free_port(), N, D, k, str(tmp_path), True), nprocs=world, join=True)
    rng = np.random.default_rng(0)
    xn = fo.normalize_rows(rng.standard_normal((N, D)).astype(np.float32))
    q = rng.standard_normal((6, D)).astype(np.float32)
    rs, ri = fo.search(xn, fo.normalize_rows(q), k)
    for r in range(world):
        g = np.load(tmp_path / f"r{r}.npz")
        assert int(g["n_bad"]) == 1 and int(g["exact_calls"]) == 1
        assert np.array_equal(g["i"], ri) and np.allclose(g["s"], rs, atol=1e-6)


def test_record_layout_alignment():
    from two_tower_model_v2_b200.sharded import record_layout, record_views
    for nq, k in [(1, 1), (3, 7), (4096, 100), (5, 1000)]:
        lay = record_layout(nq, k)
        assert lay.off_ids % 16 == 0 and lay.off_bound % 16 == 0 and lay.off_flags % 16 == 0 and lay.nbytes % 16 == 0
        assert lay.off_ids >= nq * k * 4 and lay.off_bound >= lay.off_ids + nq * k * 8
        buf = torch.zeros((2, lay.nbytes), dtype=torch.uint8)
        s, i, b, f = record_views(buf, lay, nq, k)
        assert s.shape == (2, nq, k) and i.shape == (2, nq, k) and b.shape == (2, nq) and f.shape == (2, nq)
        i[1, nq - 1, k - 1] = -1
        assert buf[1, lay.off_ids + (nq * k - 1) * 8:lay.off_ids + nq * k * 8].tolist() == [255] * 8


def test_two_phase_global_threshold_exchange(tmp_path):
    N, D, k, world = 800, 16, 10, 2
    mp.spawn(_worker, args=(world, _free_port(), N, D, k, str(tmp_path), True, True), nprocs=world, join=True)
    rng = np.random.default_rng(0)
    xn = fo.normalize_rows(rng.standard_normal((N, D)).astype(np.float32))
    q = rng.standard_normal((6, D)).astype(np.float32)
    rs, ri = fo.search(xn, fo.normalize_rows(q), k)
    for r in range(world):
        g = np.load(tmp_path / f"r{r}.npz")
        assert int(g["n_bad"]) == 1
        assert np.array_equal(g["i"], ri) and np.allclose(g["s"], rs, atol=1e-6)
