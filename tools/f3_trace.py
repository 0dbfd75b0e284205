#!/usr/bin/env python3
"""Per-phase clock64 trace of one CTA of the tcgen05 generators (QPNET_GEN_TRACE_STEP; qp_generate_f3.cu up to 128
utterances, qp_generate_f3x2.cu above or with --x2: both groups, any CTA through QPNET_GEN_TRACE_CTA).

    python tools/f3_trace.py [--utts 128] [--frames 20] [--step 1000] [--x2]

Events per phase j (block j; phase L = final skip): 0 MMA thread sees the staged z_{j-1} tile, 1 gate MMAs committed,
2 ET thread 0: tile complete in TMEM, 3 partial rows sent, 4 partial rows of the cluster arrived, 5 z_j published,
6 PZ thread 0: z_{j-1} polled and staged.  Printed per phase: the critical chain
    published(j-1) -> staged -> MMA issue -> tile in TMEM -> sent -> arrived -> published(j).
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts", type=int, default=128)
    ap.add_argument("--frames", type=int, default=20)
    ap.add_argument("--step", type=int, default=1000)
    ap.add_argument("--x2", action="store_true", help="the two-group kernel also for <= 128 utterances")
    args = ap.parse_args()
    os.environ["QPNET_GEN_TRACE_STEP"] = str(args.step)
    x2 = args.x2 or args.utts > 128
    os.environ["QPNET_GEN_KERNEL"] = "f3x2" if x2 else "f3"
    import bench
    from qpnet_b200 import _lib, ops
    from qpnet_b200.qpnet import QPNet, initialize
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    m = QPNet()
    m.apply(initialize)
    m = m.to(dev)
    L = len(m.dilationsF) + len(m.dilationsA)
    h, f0, n_list = bench.build_inputs(args.utts, 0, args.frames)
    d64, _ = ops.f0_to_dilated(torch.from_numpy(f0).to(dev), 22050, 8, 110, want_f32=False)
    seed = torch.full((args.utts,), 128, dtype=torch.int64, device=dev)
    n_dev = torch.tensor(n_list, dtype=torch.int32, device=dev)
    hd = torch.from_numpy(h).to(dev)
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        m.generate_device(seed, hd, d64, n_dev, max(n_list), check_status=False)
        e1.record()
        torch.cuda.synchronize()
        print(f"kernel ms {e0.elapsed_time(e1):.2f}  us/step {e0.elapsed_time(e1) * 1e3 / (max(n_list) + 16):.2f}")
    nphase, nev, ngt = L + 4, (64 if x2 else 24), (8 if x2 else 4)
    n_g = 128 * nphase * ngt
    n = 8 * nphase * nev + n_g
    buf = (C.c_longlong * n)()
    fn = _lib.lib.qp_debug_gen_trace
    fn.restype = C.c_int
    fn.argtypes = [C.POINTER(_lib.QpArch), C.c_int32, C.c_int32, C.c_void_p, C.c_size_t, C.POINTER(C.c_longlong), C.c_int32, C.c_void_p]
    ws = m._last_ws
    fn(m._arch, args.utts, ops.max_ceil(d64), ws.data_ptr(), ws.numel(), buf, n, None)
    allbuf = np.array(buf, dtype=np.int64)
    tr = allbuf[:8 * nphase * nev].reshape(8, nphase, nev)
    gt2 = allbuf[8 * nphase * nev:].reshape(128, nphase, ngt // 4, 4).astype(np.float64)
    gt2[gt2 == 0] = np.nan
    gt = gt2[:, :, 0, :]
    gt[gt == 0] = np.nan
    labels = ["pub(j-1)->staged", "staged->mma", "mma issue", "commit->TMEM seen", "ld+send", "send->arrived", "finish+publish"]
    acc = np.zeros(7)
    cnt = 0
    for st in range(1, 7):
        step_total = tr[st + 1, 0, 5] - tr[st, 0, 5]
        if st <= 2:
            print(f"--- step {args.step + st}: {step_total} cycles (block-0 publish to block-0 publish); block 0: symbols+tables {tr[st, 0, 5] - tr[st, 0, 2]}")
        for j in range(1, L):
            e, prev = tr[st, j], tr[st, j - 1]
            d = [e[6] - prev[5], e[0] - e[6], e[1] - e[0], e[2] - e[1], e[3] - e[2], e[4] - e[3], e[5] - e[4]]
            if st <= 2:
                print(f"blk{j:02d} " + "  ".join(f"{lab} {int(v):5d}" for lab, v in zip(labels, d)) + f"  | phase {int(e[5] - prev[5]):5d}")
            acc += d
            cnt += 1
        if st <= 2:
            print(f"final+heads+sampling+block0: {tr[st + 1, 0, 5] - tr[st, L - 1, 5]} cycles")
    # full timeline of two phases of one step, relative to the moment the MMA thread saw z_{j-1}
    names = {0: "MMA z seen", 1: "MMA z-products committed", 8: "MMA x tile seen", 9: "MMA past tile seen", 10: "MMA phase issued",
             2: "ET tile in TMEM", 3: "ET sent", 4: "ET arrived", 5: "ET z published", 11: "PZ z buffer free", 12: "PZ first piece fresh", 6: "PZ z staged",
             13: "EU tile in TMEM", 14: "EU sent", 15: "EU arrived", 16: "EU x published", 17: "PX past buffer free", 18: "PX past copies issued",
             19: "PX x buffer free", 20: "PX first x piece fresh", 21: "PX x staged"}
    names.update({22: "MMA Wc chunk seen", 23: "MMA Wp chunk seen", 24: "MMA z UMMAs issued", 25: "MMA commit 2 done", 26: "MMA commit 3 done", 27: "MMA commit 4 done", 28: "MMA x UMMAs issued", 29: "MMA x commit done"})
    for j in (5, 6, 13):
        e = tr[2, j]
        base = e[0]
        print(f"--- timeline of phase {j} (step {args.step + 2}), cycles relative to 'MMA z seen' of group 0:")
        evs = [(int(e[32 * g + k]), f"g{g} {names[k]}") for g in range(2 if x2 else 1) for k in names if 32 * g + k < nev and e[32 * g + k] != 0]
        for tstamp, label in sorted(evs):
            print(f"    {tstamp - base:8d}  {label}")
    acc /= cnt
    print("mean over blocks 1..L-1 of 6 steps: " + "  ".join(f"{lab} {v:.0f}" for lab, v in zip(labels, acc)) + f"  | phase {acc.sum():.0f}")
    # every CTA on the global clock (ns), one step: how far apart do the CTAs publish / see a complete tile?
    print(f"--- step {args.step + 2} on %globaltimer (ns; 1 ns ~ 1.9 cycles): per phase over all CTAs")
    t0 = np.nanmin(gt[:, 1, 1])
    for j in range(1, L):
        pub, stg, arr, xpub = gt[:, j, 0], gt[:, j, 1], gt[:, j, 2], gt[:, j, 3]
        nxt = gt[:, j + 1, 1] if j + 1 < L + 1 else stg
        print(f"blk{j:02d} staged {np.nanmin(stg) - t0:8.0f} .. {np.nanmax(stg) - t0:8.0f} | arrived {np.nanmin(arr) - t0:8.0f} .. {np.nanmax(arr) - t0:8.0f} | "
              f"z published {np.nanmin(pub) - t0:8.0f} .. {np.nanmax(pub) - t0:8.0f} (slowest CTA {int(np.nanargmax(pub))}) | x published {np.nanmin(xpub) - t0:8.0f} .. {np.nanmax(xpub) - t0:8.0f} | "
              f"next staged - last published {np.nanmin(nxt) - np.nanmax(pub):6.0f} .. {np.nanmax(nxt) - np.nanmax(pub):6.0f}")
    if x2 and not np.all(np.isnan(gt2[:, :, 1, :])):
        print("--- both groups, z staged (min .. max over CTAs) and z published, ns:")
        for j in range(1, L):
            print(f"blk{j:02d} " + " | ".join(f"g{g} staged {np.nanmin(gt2[:, j, g, 1]) - t0:8.0f} .. {np.nanmax(gt2[:, j, g, 1]) - t0:8.0f} published {np.nanmin(gt2[:, j, g, 0]) - t0:8.0f} .. {np.nanmax(gt2[:, j, g, 0]) - t0:8.0f}" for g in range(2)))
        # are the CTAs split into camps that serve the groups in opposite phase?  per CTA: when it staged z of block 6 for
        # each group (relative to the first CTA), and the distance between its two groups
        j = 6
        s0, s1 = gt2[:, j, 0, 1] - np.nanmin(gt2[:, j, 0, 1]), gt2[:, j, 1, 1] - np.nanmin(gt2[:, j, 0, 1])
        bins = np.arange(0, 12001, 1000)
        print("blk06 z staged, ns after the first CTA: histogram per 1 us bin")
        print("   group 0:", np.histogram(s0[~np.isnan(s0)], bins)[0].tolist())
        print("   group 1:", np.histogram(s1[~np.isnan(s1)], bins)[0].tolist())
        dd = s1 - s0
        print("   g1 - g0 per CTA: histogram per 1 us bin from -6 us:", np.histogram(dd[~np.isnan(dd)], np.arange(-6000, 6001, 1000))[0].tolist())
        late = np.argsort(-s0)[:16]
        print("   the 16 CTAs that stage group 0 last:", [(int(c), int(s0[c])) for c in late])
        for cta in (0, 1, 26, 112):
            print(f"CTA {cta}: " + "  ".join(f"j{j}: g0 {gt2[cta, j, 0, 1] - t0:.0f} g1 {gt2[cta, j, 1, 1] - t0:.0f}" for j in range(3, 8)))
    tot = np.mean([tr[st + 1, 0, 5] - tr[st, 0, 5] for st in range(1, 7)])
    tail = np.mean([tr[st + 1, 0, 5] - tr[st, L - 1, 5] for st in range(1, 7)])
    print(f"step total (mean) {tot:.0f} cycles; after the last gate (final skip, heads, sampling, block 0): {tail:.0f}")


if __name__ == "__main__":
    main()
