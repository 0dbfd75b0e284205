#!/usr/bin/env python3
"""A/B timing of the generator kernels on the bench workload shape (device-resident inputs, CUDA events).

    python tools/ab_kernels.py [--utts 32] [--frames 200] [--kernels f3,fold2]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts", type=int, default=32)
    ap.add_argument("--frames", type=int, default=200)
    ap.add_argument("--kernels", default="f3,fold2")
    ap.add_argument("--reps", type=int, default=2)
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
    ref = None
    for k in args.kernels.split(","):
        os.environ["QPNET_GEN_KERNEL"] = k
        best = 1e30
        for _ in range(args.reps + 1):
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            t0.record()
            out, _ = m.generate_device(seed, hd, d64, n_dev, max(n_list))
            t1.record()
            torch.cuda.synchronize()
            best = min(best, t0.elapsed_time(t1))
        steps = max(n_list) + 16
        tot = sum(n_list)
        same = "" if ref is None else f"  symbols equal to {args.kernels.split(',')[0]}: {float((out == ref).float().mean()):.4f}"
        if ref is None:
            ref = out
        print(f"{k:8s} {best:9.2f} ms  {best * 1e3 / steps:7.2f} us/step  {tot / best * 1e3:12.0f} samples/s  RTF {tot / best * 1e3 / 22050:6.2f}{same}", flush=True)


if __name__ == "__main__":
    main()
