#!/usr/bin/env python3
"""Bring-up aid for the tcgen05 backward (QPNET_BWD_TC bit mask, read once per process):
    python tools/bwd_tc_probe.py dump  out.pt     # gradients of the bf16 path under the current environment
    python tools/bwd_tc_probe.py cmp   ref.pt got.pt
The dump under QPNET_BWD_TC=0 (TF32 mma.sync backward on the same saved activations) is the reference: the tcgen05
kernels only add the bf16 rounding of dgate / dX / dskip, so per-tensor relative L2 differences stay around 1e-2."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CFGS = [(64, 64, 1, 12, 330, 1.0), (192, 64, 1, 12, 129, 1.5), (128, 128, 2, 14, 457, 0.5)]


def dump(path):
    from oracle import qpnet_oracle as orc
    from tests import cases
    from qpnet_b200 import synth
    from qpnet_b200.qpnet import QPNet
    dev = torch.device("cuda:0")
    out = {}
    for ci, (C, S, B, frames, bl, fac) in enumerate(CFGS):
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
        m = QPNet(**kw); m.load_state_dict(p); m.tensor_cores = True; m = m.to(dev)
        loss = torch.nn.functional.cross_entropy(m(x, h, d, blt).reshape(-1, a.Q), tgt.reshape(-1))
        loss.backward()
        torch.cuda.synchronize()
        out[ci] = {k: v.grad.detach().cpu() for k, v in m.named_parameters()}
    torch.save(out, path)
    print("dumped", path, "QPNET_BWD_TC =", os.environ.get("QPNET_BWD_TC"), "QPNET_MN_SWAP =", os.environ.get("QPNET_MN_SWAP"))


def cmp(ref, got):
    r, g = torch.load(ref), torch.load(got)
    for ci in r:
        errs = {}
        for k, v in r[ci].items():
            if v.numel() > 1 and float(v.abs().max()) > 0:
                errs[k] = float((v - g[ci][k]).norm() / v.norm())
        top = sorted(errs, key=lambda k: -errs[k])[:4]
        print(f"cfg {CFGS[ci]}: worst " + "  ".join(f"{k}={errs[k]:.4f}" for k in top))


if __name__ == "__main__":
    if sys.argv[1] == "dump":
        dump(sys.argv[2])
    else:
        cmp(sys.argv[2], sys.argv[3])
