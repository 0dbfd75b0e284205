#!/usr/bin/env python3
"""Per-tensor gradient errors of the teacher-forced pass on the bench segment (fp32 and bf16 paths) against the oracle."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import qpnet_oracle as orc
from qpnet_b200 import ops
from qpnet_b200.qpnet import QPNet
from tests.test_gpu_scale import bench_segment

torch.set_num_threads(os.cpu_count() or 1)
dev = torch.device("cuda:0")
a = orc.Arch()
p = orc.init_params(a, 12, 0.05)
x, h, d, t, bl, R = bench_segment(a, 0)
pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
want = orc.forward(a, pr, x, h, d, bl)
torch.nn.functional.cross_entropy(want.reshape(-1, a.Q), t.reshape(-1)).backward()
# the oracle against itself in float64: how much of the difference is the oracle's own fp32 rounding?
p64 = {k: v.double().clone().requires_grad_(True) for k, v in p.items()}
w64 = orc.forward(a, p64, x, h.double(), d, bl) if os.environ.get("DIAG_F64") else None
if w64 is not None:
    torch.nn.functional.cross_entropy(w64.reshape(-1, a.Q), t.reshape(-1)).backward()
for tc in (False, True):
    m = QPNet(); m.load_state_dict(p); m = m.to(dev); m.tensor_cores = tc
    got = m(x.to(dev), h.to(dev), d.to(dev), torch.tensor([bl], device=dev))
    loss, dl = ops.cross_entropy(got.detach(), t.to(dev))
    got.backward(dl)
    rows = []
    for k, prm in m.named_parameters():
        ref = pr[k].grad
        if ref is None or float(ref.abs().max()) == 0:
            continue
        gg = prm.grad.detach().cpu()
        e_max = float((gg - ref).abs().max() / ref.abs().max())
        e_l2 = float((gg - ref).norm() / ref.norm())
        o64 = ""
        if w64 is not None and p64[k].grad is not None:
            r64 = p64[k].grad
            o64 = " | oracle fp32 vs fp64: max %.2e l2 %.2e ; ours vs fp64: max %.2e" % (
                float((ref.double() - r64).abs().max() / r64.abs().max()), float((ref.double() - r64).norm() / r64.norm()),
                float((gg.double() - r64).abs().max() / r64.abs().max()))
        rows.append((e_max, e_l2, k, o64))
    rows.sort(reverse=True)
    print(f"== tensor_cores={tc}: max |dlogit| {float((got.detach().cpu() - want.detach()).abs().max()):.3e}")
    for e_max, e_l2, k, o64 in rows[:12]:
        print(f"  {k:32s} max/max {e_max:.3e}  rel L2 {e_l2:.3e}{o64}")
    print("  median rel L2 %.3e" % float(np.median([r[1] for r in rows])))
