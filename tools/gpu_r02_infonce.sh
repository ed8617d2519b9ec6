#!/bin/bash
timeout 300 python -m pytest -q -p no:cacheprovider --timeout=200 tests/test_gpu_infonce.py 2>&1 | tail -12
timeout 120 python - <<'PY'
import torch, time
import two_tower_model_v2_b200 as pkg
import sys; sys.path.insert(0, '.')
from oracle import infonce_oracle as io
B, M, D = 512, 4, 384
g = torch.Generator(device="cuda").manual_seed(1)
b = torch.nn.functional.normalize(torch.randn(B, D, device="cuda", generator=g), dim=1).requires_grad_(True)
p = torch.nn.functional.normalize(torch.randn(B, D, device="cuda", generator=g), dim=1).requires_grad_(True)
n = torch.nn.functional.normalize(torch.randn(B, M, D, device="cuda", generator=g), dim=2).requires_grad_(True)
crit = pkg.InfoNCELoss()
def ours():
    l = crit(b, p, n); l.backward(); return l
def ref():
    l = io.torch_loss(b, p, n, 0.07); l.backward(); return l
for f, name in ((ours, "tt kernels fwd+bwd"), (ref, "torch (reference arithmetic) fwd+bwd")):
    for _ in range(5): f()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
    for _ in range(50): f()
    e1.record(); torch.cuda.synchronize()
    print(name, round(e0.elapsed_time(e1) / 50 * 1e3, 1), "us per step at B=512, M=4, D=384")
PY
