"""GPU: the InfoNCE kernels (tt_infonce_forward / tt_infonce_backward, through the C-ABI) against the golden vectors of
the real reference `InfoNCELoss` (src/training/losses.py:36-79), the oracle and fp64 autograd of its arithmetic.
Tolerance: 1e-5 relative on the loss, 1e-5 of the largest gradient element on every gradient."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN, golden
from oracle import infonce_oracle as io

pytestmark = pytest.mark.gpu

CASES = sorted(p.name for p in GOLDEN.glob("infonce_*.npz"))


def _close(a, ref, tol=1e-5):
    a, ref = np.asarray(a, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    assert a.shape == ref.shape
    if ref.size:
        assert np.abs(a - ref).max() <= tol * max(np.abs(ref).max(), 1e-30), (np.abs(a - ref).max(), np.abs(ref).max())


@pytest.mark.parametrize("case", CASES)
def test_infonce_matches_reference_golden(case):
    from two_tower_model_v2_b200 import InfoNCELoss
    g = golden(case)
    t = float(g["temperature"])
    b, p, n = (torch.from_numpy(g[k]).cuda().requires_grad_(True) for k in ("buyer", "pos", "neg"))
    loss = InfoNCELoss(t)(b, p, n)
    assert loss.shape == () and abs(loss.item() - float(g["loss"])) <= 1e-5 * max(1.0, abs(float(g["loss"])))
    loss.backward()
    _close(b.grad.cpu().numpy(), g["d_buyer"])
    _close(p.grad.cpu().numpy(), g["d_pos"])
    _close(n.grad.cpu().numpy(), g["d_neg"])


def test_infonce_training_batch_against_fp64_autograd():
    """Trainer shape (batch 512, 4 sampled negatives, D = 384; config.yaml training.batch_size / trainer.py:216-236) with
    an upstream gradient != 1, against fp64 autograd of the reference arithmetic and against the row-wise oracle."""
    from two_tower_model_v2_b200 import InfoNCELoss, ops
    rng = np.random.default_rng(5)
    B, M, D = 512, 4, 384
    b = rng.standard_normal((B, D)).astype(np.float32)
    p = (0.6 * b + 0.8 * rng.standard_normal((B, D))).astype(np.float32)       # positives correlate with their buyers
    n = rng.standard_normal((B, M, D)).astype(np.float32)
    b, p = b / np.linalg.norm(b, axis=1, keepdims=True), p / np.linalg.norm(p, axis=1, keepdims=True)
    n = n / np.linalg.norm(n, axis=2, keepdims=True)
    tb, tp, tn = (torch.from_numpy(a).cuda().requires_grad_(True) for a in (b, p, n))
    loss = InfoNCELoss(0.07)(tb, tp, tn)
    (3.0 * loss).backward()
    rb, rp, rn = (torch.from_numpy(a).double().requires_grad_(True) for a in (b, p, n))
    ref = io.torch_loss(rb, rp, rn, 0.07)
    (3.0 * ref).backward()
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
    _close(tb.grad.cpu().numpy(), rb.grad.numpy())
    _close(tp.grad.cpu().numpy(), rp.grad.numpy())
    _close(tn.grad.cpu().numpy(), rn.grad.numpy())
    # the per-row outputs of the C entry point
    l1, row, lse = ops.infonce_forward(tb.detach(), tp.detach(), tn.detach(), 0.07)
    o_loss, o_row, o_lse = io.loss(b.astype(np.float64), p.astype(np.float64), n.astype(np.float64), 0.07)
    _close(row.cpu().numpy(), o_row)
    _close(lse.cpu().numpy(), o_lse)
    assert abs(l1.item() - o_loss) <= 1e-5 * abs(o_loss)


def test_infonce_argument_errors():
    from two_tower_model_v2_b200 import _native
    lib = _native.load()
    x = torch.zeros(8, device="cuda")
    assert lib.tt_infonce_forward(x.data_ptr(), x.data_ptr(), 0, 2, 1, 4, 0.07, x.data_ptr(), x.data_ptr(), x.data_ptr(), 0) != 0
    assert b"NULL" in lib.tt_last_error()
    assert lib.tt_infonce_forward(x.data_ptr(), x.data_ptr(), 0, 2, 0, 2000, 0.07, x.data_ptr(), x.data_ptr(), x.data_ptr(), 0) != 0
    assert lib.tt_infonce_forward(x.data_ptr(), x.data_ptr(), 0, 2, 0, 4, 0.0, x.data_ptr(), x.data_ptr(), x.data_ptr(), 0) != 0
