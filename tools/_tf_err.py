import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import qpnet_oracle as orc
from tests import cases
from qpnet_b200.qpnet import QPNet
name = "full_sampling"
g = cases.load("generate")
kw, a, p, x, h, d, n_list, mode, xm = cases.generate_inputs(name)
dev = torch.device("cuda:0")
m = QPNet(**kw); m.load_state_dict(p); m = m.to(dev)
B = len(n_list); steps = 60
forced = torch.stack([torch.from_numpy(g[f"{name}/sym{b}"][:steps].astype(np.int64)) for b in range(B)])
lg = []
with torch.no_grad():
    orc.generate(a, p, x, h, list(n_list), d, mode="argmax", force=forced, logits_out=lg, max_steps=steps)
want = torch.stack(lg, dim=1)
res, got = m.batch_fast_generate(x, h, [steps] * B, d, None, "argmax", False, force=forced, return_logits=True)
err = (got.cpu() - want).abs().amax(dim=2)
print("per-step max |dlogit| utt0:", np.round(err[0, :24].numpy(), 4))
print("per-step max |dlogit| utt1:", np.round(err[1, :24].numpy(), 4))
print("overall", float(err.max()))
