#!/usr/bin/env python3
"""Where a training step's time goes (CUDA events around forward / CE / backward / Adam)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qpnet_b200 import ops, synth
from qpnet_b200.qpnet import QPNet, initialize
from qpnet_b200.train import Trainer, segment_geometry
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = QPNet(); model.apply(initialize); model = model.to(dev); model.check_range = False
tr = Trainer(model, lr=1e-4)
hs, f0, _ = synth.utterance(260, 700)
d64, d32 = ops.f0_to_dilated(torch.from_numpy(f0[None]).to(dev), synth.FS, synth.DENSE_FACTOR, synth.UPSAMPLING)
R, bl, h_bs, x_bs = segment_geometry(float(d32.max()), 20000, synth.UPSAMPLING, model.receptiveCausal_field, model.receptiveF_field, model.receptiveA_field)
wav = synth.noise_waveform(x_bs, 0)
xq = ops.mulaw_encode_t(torch.from_numpy(wav.astype(np.float64)).to(dev))
x, t = xq[None, :-1].contiguous(), xq[None, 1:].contiguous()
h = torch.from_numpy(hs[:h_bs].T.copy())[None].to(dev)
d = d32[:, : x_bs - 1].contiguous()
blt = torch.full((1,), bl, dtype=torch.long, device=dev)
def ev(): return torch.cuda.Event(enable_timing=True)
for tc in (True, False):
    model.tensor_cores = tc
    for _ in range(2): tr.step(x, h, d, t, bl)
    acc = np.zeros(4)
    for _ in range(3):
        e = [ev() for _ in range(5)]
        tr.optimizer.zero_grad(set_to_none=True)
        e[0].record(); logits = model(x, h, d, blt)
        e[1].record(); loss, dl = ops.cross_entropy(logits.detach(), t[:, -bl:])
        e[2].record(); logits.backward(dl)
        e[3].record(); tr.optimizer.step()
        e[4].record(); torch.cuda.synchronize()
        acc += [e[i].elapsed_time(e[i + 1]) for i in range(4)]
    acc /= 3
    print(("bf16 tcgen05 forward" if tc else "fp32 forward"), "ms: forward %.1f  CE %.2f  backward %.1f  adam %.2f  total %.1f" % (*acc, acc.sum()))
from qpnet_b200 import _lib
print("weight-gradient operand segments bound to TMA descriptors so far:", _lib.lib.qp_debug_tma_segments())
