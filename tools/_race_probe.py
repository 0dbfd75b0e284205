import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from tests import cases
from oracle import qpnet_oracle as orc
from qpnet_b200.qpnet import QPNet
dev = torch.device("cuda:0")
g = cases.load("generate")
name = "full_sampling"
kw, a, p, x, h, d, n_list, mode, xm = cases.generate_inputs(name)
m = QPNet(**kw); m.load_state_dict(p); m = m.to(dev)
steps = int(os.environ.get("STEPS", "24"))
B = len(n_list)
forced = torch.stack([torch.from_numpy(g[f"{name}/sym{b}"][:steps].astype(np.int64)) for b in range(B)])
lg = []
with torch.no_grad():
    orc.generate(a, p, x, h, list(n_list), d, mode="argmax", force=forced, logits_out=lg, max_steps=steps)
want = torch.stack(lg, dim=1)
for rep in range(int(os.environ.get("REPS", "3"))):
    res, got = m.batch_fast_generate(x, h, [steps] * B, d, None, "argmax", xm, force=forced, return_logits=True)
    e = (got.cpu() - want).abs().amax(dim=2)
    print("rep", rep, "max err", float(e.max()), "first bad step per utt", [int((e[b] > 0.06).nonzero()[0]) if (e[b] > 0.06).any() else -1 for b in range(B)], flush=True)
