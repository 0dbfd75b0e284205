import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
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
tr.step(x, h, d, t, bl); torch.cuda.synchronize()
