#!/usr/bin/env python3
"""One small batch_fast_generate call (profiling harness: ncu wraps this, never bench.py).

    python tools/gen_once.py [--utts 32] [--frames 5] [--reps 1]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts", type=int, default=32)
    ap.add_argument("--frames", type=int, default=5)
    ap.add_argument("--reps", type=int, default=1)
    args = ap.parse_args()
    import bench
    from qpnet_b200 import ops
    from qpnet_b200.qpnet import QPNet, initialize
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    m = QPNet()
    m.apply(initialize)
    m = m.to(dev)
    h, f0, n_list = bench.build_inputs(args.utts, 0, args.frames)
    d64, _ = ops.f0_to_dilated(torch.from_numpy(f0).to(dev), 22050, 8, 110, want_f32=False)
    seed = torch.full((args.utts,), 128, dtype=torch.int64, device=dev)
    n_dev = torch.tensor(n_list, dtype=torch.int32, device=dev)
    hd = torch.from_numpy(h).to(dev)
    for _ in range(args.reps):
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        out, _ = m.generate_device(seed, hd, d64, n_dev, max(n_list), n_host=n_list)
        t1.record()
        torch.cuda.synchronize()
        print("ms", t0.elapsed_time(t1), "us/step", t0.elapsed_time(t1) * 1e3 / (max(n_list) + 1), "sym[0,:8]", out[0, :8].tolist())


if __name__ == "__main__":
    main()
