import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import qpnet_oracle as orc
from tests import cases
from qpnet_b200 import synth
from qpnet_b200.qpnet import QPNet
dev = torch.device("cuda:0")
def run(C, S, B, frames, bl, fac):
    kw = dict(n_resch=C, n_skipch=S)
    a = orc.Arch(**kw); p = orc.init_params(a, 31, 0.1)
    T = frames * a.U
    xs, hs, ds = [], [], []
    for b in range(B):
        hh, f0, _ = synth.utterance(frames, 40 + b, fac, a.A)
        ds.append(torch.from_numpy(cases.d_from_f0(f0)).float()[:T]); hs.append(torch.from_numpy(hh.T.copy()))
        xs.append(torch.from_numpy(np.random.RandomState(b).randint(0, a.Q, size=T)).long())
    x, h, d = torch.stack(xs).to(dev), torch.stack(hs).to(dev), torch.stack(ds).to(dev)
    blt = torch.tensor([bl] * B, device=dev)
    tgt = torch.from_numpy(np.random.RandomState(7).randint(0, a.Q, size=(B, bl))).long().to(dev)
    pq = {k: (v.to(torch.bfloat16).float() if v.dim() > 1 else v) for k, v in p.items()}     # weights rounded to bf16
    grads = []
    for tc, prm in ((False, p), (True, p), (False, pq)):
        m = QPNet(**kw); m.load_state_dict(prm); m.tensor_cores = tc; m = m.to(dev)
        loss = torch.nn.functional.cross_entropy(m(x, h, d, blt).reshape(-1, a.Q), tgt.reshape(-1))
        loss.backward()
        grads.append({k: v.grad.clone() for k, v in m.named_parameters()})
    def rel(i):
        out = {}
        for k in grads[0]:
            r = grads[0][k]
            if r.numel() > 1 and float(r.abs().max()) > 0:
                out[k] = float((r - grads[i][k]).norm() / r.norm())
        return out
    e_tc, e_q = rel(1), rel(2)
    top = sorted(e_tc, key=lambda k: -e_tc[k])[:5]
    print(f"C={C} S={S} B={B} fac={fac}:")
    for k in top:
        print(f"   {k:28s} tcgen05-forward {e_tc[k]:.3f}   fp32 path with bf16-rounded weights {e_q[k]:.3f}")
for cfg in [(64, 64, 1, 12, 330, 1.0), (192, 64, 1, 12, 129, 1.5), (128, 64, 2, 14, 457, 1.0)]:
    run(*cfg)
