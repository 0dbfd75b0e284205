#!/usr/bin/env python3
"""Training-step benchmark: BASELINE.json configs[4] (SI-QPNet training step on synthetic noise-shaped
waveform segments + aux features, data parallel with an NCCL gradient all-reduce).

    python tools/bench_train.py [--steps K] [--warmup W] [--fp32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_train.py --gpus N

One "step" = forward (QP_F_SAVE) + fused softmax-CE + hand-written backward + gradient all-reduce + Adam on one
segment per rank (the reference's batch_size 1 per GPU, qpnet_train.py:416-423, param_model.py:58-64).
Prints ONE JSON line: train seg/s (whole job) and the achieved fraction of the measured bf16 tensor peak
from the algorithmic 141.5 MFLOP per sample.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def measure(dev, rank, world, steps=5, warmup=3, fp32=False, batch_length=20000):
    """Time `steps` training steps on every rank (device-timed, max over ranks); rank 0 gets the result dict."""
    from qpnet_b200 import ops, synth
    from qpnet_b200.qpnet import QPNet, initialize
    from qpnet_b200.train import Trainer, segment_geometry

    torch.manual_seed(0)
    model = QPNet()
    model.apply(initialize)
    model = model.to(dev)
    model.tensor_cores = not fp32
    model.check_range = False
    tr = Trainer(model, lr=1e-4)

    # one synthetic segment per rank, cut the reference's way (qpnet_train.py:268-303)
    frames = 260
    hs, f0, _ = synth.utterance(frames, 700 + rank)
    d64, d32 = ops.f0_to_dilated(torch.from_numpy(f0[None]).to(dev), synth.FS, synth.DENSE_FACTOR, synth.UPSAMPLING)
    R, bl, h_bs, x_bs = segment_geometry(float(d32.max()), batch_length, synth.UPSAMPLING,
                                         model.receptiveCausal_field, model.receptiveF_field, model.receptiveA_field)
    wav = synth.noise_waveform(x_bs, rank)
    xq = ops.mulaw_encode_t(torch.from_numpy(wav.astype(np.float64)).to(dev))        # (x_bs,) int64
    x, t = xq[None, :-1].contiguous(), xq[None, 1:].contiguous()
    h = torch.from_numpy(hs[:h_bs].T.copy())[None].to(dev)
    d = d32[:, : x_bs - 1].contiguous()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    losses = []
    for _ in range(warmup):
        losses.append(float(tr.step(x, h, d, t, bl)))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = tr.step(x, h, d, t, bl)
    e1.record()
    barrier()
    losses.append(float(loss))
    tdev = torch.tensor([e0.elapsed_time(e1) / 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tdev, op=dist.ReduceOp.MAX)
    sec = float(tdev.item())
    if rank != 0:
        return None
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = float(peaks.get("bf16_tflops_sustained", 1400.0)) * world      # whole-job FLOP/s against world x one GPU's peak
    seg_s = world * steps / sec
    flops = seg_s * bl * 141.5e6
    return {"metric": "train seg/s", "value": seg_s, "unit": "segments/s", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": sec / steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "dtype": "fp32" if fp32 else model.train_dtype,
            "data": "synthetic",
            "config": {"workload": "BASELINE configs[4]: SI-QPNet training step, 1 segment per rank", "bl": bl,
                       "receptive_field": R, "segment_samples": x_bs - 1, "parallelism": f"dp{world}, one NCCL all-reduce of {tr.bucket.numel} fp32 gradients"},
            "roofline": {"bound": "tensor", "achieved": flops / 1e12, "peak": peak, "peak_is": f"{world} x measured sustained bf16 (MEASURED_PEAKS.json)", "unit": "TFLOP/s", "frac": flops / 1e12 / peak,
                         "algorithmic_flops_per_step": bl * 141.5e6},
            "loss_first_last": [losses[0], losses[-1]], "gpu_launches_per_step": model.last_launches}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--fp32", action="store_true", help="exact fp32 SIMT forward instead of the bf16 tcgen05 path")
    ap.add_argument("--batch-length", type=int, default=20000)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    import __graft_entry__
    if rank == 0:
        __graft_entry__.build()
    if world > 1:
        dist.barrier()
    res = measure(dev, rank, world, args.steps, args.warmup, args.fp32, args.batch_length)
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
