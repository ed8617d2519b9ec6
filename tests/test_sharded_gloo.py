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
    """Stands in for FlatIPIndex on a CPU rank: same search_checked_device contract."""

    def __init__(self, xn_shard, id_offset):
        self.xn, self.id_offset = xn_shard, id_offset

    @property
    def ntotal(self):
        return self.xn.shape[0]

    def search_checked_device(self, q, k):
        s, i = fo.search(self.xn, fo.normalize_rows(q.numpy()), k)
        return torch.from_numpy(s), torch.from_numpy(i + self.id_offset), 0


def _merge_oracle(sg, ig):
    G, nq, K = sg.shape
    s = sg.permute(1, 0, 2).reshape(nq, G * K).numpy()
    i = ig.permute(1, 0, 2).reshape(nq, G * K).numpy()
    out_s, out_i = np.empty((nq, K), np.float32), np.empty((nq, K), np.int64)
    for r in range(nq):
        order = np.lexsort((i[r], -s[r]))[:K]
        out_s[r], out_i[r] = s[r][order], i[r][order]
    return torch.from_numpy(out_s), torch.from_numpy(out_i)


def _worker(rank, world, port, N, D, k, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import two_tower_model_v2_b200 as pkg
    rng = np.random.default_rng(0)                      # same catalog and queries on every rank
    xn = fo.normalize_rows(rng.standard_normal((N, D)).astype(np.float32))
    q = rng.standard_normal((6, D)).astype(np.float32)
    lo, hi = pkg.shard_bounds(N, world, rank)
    sharded = pkg.ShardedFlatIPIndex(_LocalOracleIndex(xn[lo:hi], lo), N, merge=_merge_oracle)
    s, i, _ = sharded.search_device(torch.from_numpy(q), k)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), s=s.numpy(), i=i.numpy())
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
