"""Minimal numpy stand-in for the `faiss` module — TEST INFRASTRUCTURE ONLY.

Exists so that the UNMODIFIED reference file /root/reference/src/inference/vector_db.py (which does
`import faiss` at line 3) can be imported in the authoring container, where faiss-cpu is not
installed, to generate golden vectors of its wrapper logic (tests/golden/make_golden.py).  It
implements only what that file calls: IndexFlatIP(d).add/.search/.ntotal, read_index, write_index.
Never on the product path; never shipped as a faiss replacement.
"""
import numpy as np

from oracle import flat_ip_oracle as _o

__version__ = "shim-0 (numpy restatement; not faiss)"

# "numpy": the tie-ordered oracle search (golden vectors, tests).  "torch": blocked sgemm + top-k on all host
# threads, what faiss-cpu does for nq >= 20 - the timed CPU baseline of bench.py sets this.
SEARCH_IMPL = "numpy"


class IndexFlatIP:
    def __init__(self, d):
        self.d = int(d)
        self._x = np.zeros((0, self.d), np.float32)

    @property
    def ntotal(self):
        return self._x.shape[0]

    def add(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2 and x.shape[1] == self.d
        self._x = np.concatenate([self._x, x], 0)

    def search(self, q, k):
        q = np.ascontiguousarray(q, dtype=np.float32)
        if SEARCH_IMPL == "torch":
            import torch
            ts, ti = _o.torch_search(torch.from_numpy(self._x), torch.from_numpy(q), min(k, self.ntotal), block=65536)
            s, i = ts.numpy(), ti.numpy()
        else:
            s, i = _o.search(self._x, q, k)
        if s.shape[1] < k:   # faiss pads missing results with -1 labels
            pad = k - s.shape[1]
            s = np.concatenate([s, np.full((s.shape[0], pad), -np.inf, np.float32)], 1)
            i = np.concatenate([i, np.full((i.shape[0], pad), -1, np.int64)], 1)
        return s, i


def write_index(index, path):
    np.save(path if str(path).endswith(".npy") else str(path) + ".npy", index._x)


def read_index(path):
    x = np.load(path if str(path).endswith(".npy") else str(path) + ".npy")
    idx = IndexFlatIP(x.shape[1])
    idx.add(x)
    return idx
