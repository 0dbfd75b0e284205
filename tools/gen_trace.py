#!/usr/bin/env python3
"""Debug aid: per-phase clock64 trace of CTA 0 of the persistent generator (QPNET_GEN_TRACE_STEP).

    python tools/gen_trace.py [--utts 32] [--frames 20] [--step 1000]
Prints, for 8 consecutive sample steps, the cycles spent per phase split into
wait (poll for inputs), mma, epilogue.
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def fold2_trace(m, args, d64, L):
    """role-split two-level folded generator (qp_generate_fold2.cu): phases = block 0, 512-phases 1..L (L = final skip),
    head-1, head-2; 13 events each, see TRACE_EVENTS in the kernel (0-7 finisher thread 0, 8-12 streamer thread 128)"""
    from qpnet_b200 import _lib, ops
    nphase, nev = L + 4, 13
    n = 8 * nphase * nev
    buf = (C.c_longlong * n)()
    fn = _lib.lib.qp_debug_gen_trace
    fn.restype = C.c_int
    fn.argtypes = [C.POINTER(_lib.QpArch), C.c_int32, C.c_int32, C.c_void_p, C.c_size_t, C.POINTER(C.c_longlong),
                   C.c_int32, C.c_void_p]
    ws = m._last_ws
    M = ops.max_ceil(d64)
    fn(m._arch, args.utts, M, ws.data_ptr(), ws.numel(), buf, n, None)
    tr = np.array(buf, dtype=np.int64).reshape(8, nphase, nev)
    names = ["blk0  "] + [f"fuse{j:02d}" for j in range(1, L)] + ["final ", "head1 ", "head2 "]
    labels = ["poll", "bar", "gateMMA", "send", "res+H+wait", "pubZ", "rest"]
    for st in range(1, 3):
        print(f"--- step {args.step + st}: total {tr[st + 1, 0, 0] - tr[st, 0, 0]} cycles")
        for ph in range(nphase - 1):
            e = tr[st, ph]
            nxt = tr[st, ph + 1, 0] if ph + 1 < nphase - 1 else tr[st + 1, 0, 0]
            if ph == 0:
                print(f"{names[ph]} symbols {e[1] - e[0]:6d}  pubZ {e[6] - e[1]:6d}  tables {e[7] - e[6]:5d}  gap {nxt - e[7]:5d}")
            elif ph < L:
                print(names[ph] + " " + "  ".join(f"{lab} {e[i + 1] - e[i]:5d}" for i, lab in enumerate(labels)) + f"  gap {nxt - e[7]:5d}"
                      + f"  | streamer vs finisher start: staged {e[8] - e[0]:6d} Esent {e[9] - e[0]:6d} xpub {e[10] - e[0]:6d} issued {e[11] - e[0]:6d} done {e[12] - e[0]:6d}")
            elif ph == L:
                print(f"{names[ph]} poll {e[1] - e[0]:5d}  bar {e[2] - e[1]:5d}  rest {e[7] - e[2]:5d}  gap {nxt - e[7]:5d}"
                      + f"  | streamer: staged {e[8] - e[0]:6d} Esent {e[9] - e[0]:6d} skip pub {e[10] - e[0]:6d} done {e[12] - e[0]:6d}")
            else:
                print(f"{names[ph]} poll {e[1] - e[0]:5d}  bar {e[2] - e[1]:5d}  mma {e[3] - e[2]:5d}  send {e[4] - e[3]:5d}  finish {e[7] - e[4]:5d}  gap {nxt - e[7]:5d}")
    crit = np.zeros(8)
    strm = np.zeros(5)
    for st in range(1, 7):
        for ph in range(1, L):
            e = tr[st, ph]
            crit[:7] += [e[i + 1] - e[i] for i in range(7)]
            crit[7] += tr[st, ph + 1, 0] - e[7]
            strm += [e[9] - e[8], e[10] - e[9], e[11] - e[10], e[12] - e[11], (tr[st, ph + 1, 8] - e[12]) if ph + 1 <= L else 0]
    crit /= 6 * (L - 1)
    strm /= 6 * (L - 1)
    print("finisher, fused phases, mean over 6 steps:", "  ".join(f"{lab} {crit[i]:.0f}" for i, lab in enumerate(labels + ["gap"])),
          " total/phase %.0f" % crit.sum())
    print("streamer, fused phases: staged->Esent %.0f  ->x published %.0f  ->tile requested %.0f  ->ring stored %.0f  ->next staged %.0f" % tuple(strm))
    tot = [tr[st + 1, 0, 0] - tr[st, 0, 0] for st in range(1, 7)]
    tail = [tr[st + 1, 0, 0] - tr[st, L, 0] for st in range(1, 7)]
    head = [tr[st, 1, 0] - tr[st, 0, 0] for st in range(1, 7)]
    print("step total (mean):", np.mean(tot), " block 0 incl. symbol wait:", np.mean(head), " final+heads:", np.mean(tail))


def fold_trace(m, args, d64, L):
    """folded generator: phases = block 0, fused 1..L-1, final skip, head-1, head-2, sampling; 8 events each:
    0 start, 1 own pieces fresh, 2 barrier passed, 3 MMA done, 4 sent, 5 x published, 6 gate partials arrived,
    7 z (or skip / head row) published"""
    from qpnet_b200 import _lib, ops
    nphase, nev = L + 4, 14
    n = 8 * nphase * nev
    buf = (C.c_longlong * n)()
    fn = _lib.lib.qp_debug_gen_trace
    fn.restype = C.c_int
    fn.argtypes = [C.POINTER(_lib.QpArch), C.c_int32, C.c_int32, C.c_void_p, C.c_size_t, C.POINTER(C.c_longlong),
                   C.c_int32, C.c_void_p]
    ws = m._last_ws
    M = ops.max_ceil(d64)
    fn(m._arch, args.utts, M, ws.data_ptr(), ws.numel(), buf, n, None)
    tr = np.array(buf, dtype=np.int64).reshape(8, nphase, nev)
    names = ["blk0  "] + [f"fuse{j:02d}" for j in range(1, L)] + ["final ", "head1 ", "head2 ", "sample"]
    labels = ["poll", "barrier", "mma", "send", "resX", "dsmemG", "pubZ"]
    for st in range(1, 3):
        print(f"--- step {args.step + st}: total {tr[st + 1, 0, 0] - tr[st, 0, 0]} cycles")
        for ph in range(nphase - 1):
            e = tr[st, ph]
            nxt = tr[st, ph + 1, 0]
            if ph == 0:
                print(f"{names[ph]} symbols {e[1] - e[0]:6d}  finish {e[6] - e[2]:6d}  pubX {e[7] - e[6]:5d}  gap {nxt - e[7]:5d}")
            elif ph < L:
                print(names[ph] + " " + "  ".join(f"{lab} {e[i + 1] - e[i]:5d}" for i, lab in enumerate(labels)) + f"  gap {nxt - e[7]:5d}"
                      + f"   | streamer (vs phase start): fresh {e[8] - e[0]:5d} tile {e[9] - e[0]:5d} [barrier {e[2] - e[0]:5d}] sent {e[10] - e[0]:5d} issued {e[12] - e[0]:5d} ringmma {e[13] - e[0]:5d} done {e[11] - e[0]:5d}")
            elif ph == L:
                print(f"{names[ph]} poll {e[1] - e[0]:5d}  barrier {e[2] - e[1]:5d}  res+finish {e[7] - e[2]:5d}  gap {nxt - e[7]:5d}")
            else:
                print(f"{names[ph]} poll {e[1] - e[0]:5d}  barrier {e[2] - e[1]:5d}  mma {e[3] - e[2]:5d}  send {e[4] - e[3]:5d}  finish {e[7] - e[4]:5d}  gap {nxt - e[7]:5d}")
    acc = np.zeros((nphase, 8))
    for st in range(1, 7):
        for ph in range(1, L):
            e = tr[st, ph]
            acc[ph, :7] += [e[i + 1] - e[i] for i in range(7)]
            acc[ph, 7] += tr[st, ph + 1, 0] - e[7]
    acc /= 6
    print("fused phases, mean over 6 steps:", "  ".join(f"{lab} {acc[1:L, i].mean():.0f}" for i, lab in enumerate(labels + ["gap"])),
          " total/phase %.0f" % acc[1:L].sum(1).mean())
    # per-CTA sums over 256 steps (tid 0): wait = phase start -> barrier, work = barrier -> z published (includes the two
    # waits for the cluster's partial tiles, dsR / dsG), sym = wait for the fed-back symbol
    sb = (C.c_longlong * (128 * 8))()
    fs = _lib.lib.qp_debug_gen_stats
    fs.restype = C.c_int
    fs.argtypes = fn.argtypes
    if fs(m._arch, args.utts, M, ws.data_ptr(), ws.numel(), sb, 128 * 8, None) > 0:
        stt = np.array(sb, dtype=np.int64).reshape(128, 8)
        per = stt[:, :5] / (256.0 * np.array([L - 1, L - 1, L - 1, L - 1, 1]))
        print("per-CTA mean cycles per fused phase: wait / work / dsR / dsG ; symbol wait per step ; smid")
        for c in range(128):
            print(f"cta {c:3d} rank {c % 4} smid {stt[c, 5]:3d}: wait {per[c, 0]:6.0f} work {per[c, 1]:6.0f} dsR {per[c, 2]:6.0f} dsG {per[c, 3]:6.0f} sym {per[c, 4]:6.0f}")
        print("by rank: " + "  ".join(f"r{r}: wait {per[r::4, 0].mean():.0f} work {per[r::4, 1].mean():.0f} dsR {per[r::4, 2].mean():.0f} dsG {per[r::4, 3].mean():.0f}" for r in range(4)))
    tot = [tr[st + 1, 0, 0] - tr[st, 0, 0] for st in range(1, 7)]
    tail = [tr[st + 1, 0, 0] - tr[st, L, 0] for st in range(1, 7)]
    print("step total (mean):", np.mean(tot), " of which final+head+sampling:", np.mean(tail))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts", type=int, default=32)
    ap.add_argument("--frames", type=int, default=20)
    ap.add_argument("--step", type=int, default=1000)
    args = ap.parse_args()
    os.environ["QPNET_GEN_TRACE_STEP"] = str(args.step)
    import bench
    from qpnet_b200 import _lib, ops
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
    for _ in range(2):
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        out, _ = m.generate_device(seed, torch.from_numpy(h).to(dev), d64, n_dev, max(n_list), check_status=False)
        t1.record()
        torch.cuda.synchronize()
        print("kernel ms", t0.elapsed_time(t1), "us/step", t0.elapsed_time(t1) * 1e3 / (max(n_list) + 1))
    L = 16
    kind = os.environ.get("QPNET_GEN_KERNEL", "fold2")
    if kind == "fold2":
        return fold2_trace(m, args, d64, L)
    if kind == "fold":
        return fold_trace(m, args, d64, L)
    nphase = 2 * L + 3
    cluster = kind != "generic"
    nev = 8 if cluster else 4
    n = 8 * nphase * nev
    buf = (C.c_longlong * n)()
    fn = _lib.lib.qp_debug_gen_trace
    fn.restype = C.c_int
    fn.argtypes = [C.POINTER(_lib.QpArch), C.c_int32, C.c_int32, C.c_void_p, C.c_size_t, C.POINTER(C.c_longlong),
                   C.c_int32, C.c_void_p]
    ws = m._last_ws
    M = ops.max_ceil(d64)
    fn(m._arch, args.utts, M, ws.data_ptr(), ws.numel(), buf, n, None)
    tr = np.array(buf, dtype=np.int64).reshape(8, nphase, nev)
    names = [f"{'gate' if i % 2 == 0 else 'res '}{i // 2:02d}" for i in range(2 * L)] + ["head1 ", "head2 ", "sample"]
    if cluster:
        # events: 0 start, 1 own pieces fresh, 2 barrier passed, 3 MMA done, 4 sent, 5 partials arrived, 6 published
        labels = ["poll", "barrier", "mma", "send", "dsmem", "finish"]
        for st in range(1, 3):
            print(f"--- step {args.step + st}: total {tr[st + 1, 0, 0] - tr[st, 0, 0]} cycles")
            for ph in range(nphase - 1):
                e = tr[st, ph]
                nxt = tr[st, ph + 1, 0]
                if ph == 0:
                    print(f"{names[ph]} symbols {e[1] - e[0]:6d}  finish {e[6] - e[2]:6d}  gap {nxt - e[6]:5d}")
                else:
                    print(names[ph] + " " + "  ".join(f"{lab} {e[i + 1] - e[i]:5d}" for i, lab in enumerate(labels)) + f"  gap {nxt - e[6]:5d}")
        acc = np.zeros((nphase, 7))
        for st in range(1, 7):
            for ph in range(1, nphase - 1):
                e = tr[st, ph]
                acc[ph, :6] += [e[i + 1] - e[i] for i in range(6)]
                acc[ph, 6] += tr[st, ph + 1, 0] - e[6]
        acc /= 6
        print("mean over 6 steps, all MMA phases:", "  ".join(f"{lab} {acc[1:-1, i].mean():.0f}" for i, lab in enumerate(labels + ["gap"])))
        gate = [2 * l for l in range(1, L)]
        res = [2 * l + 1 for l in range(L)]
        print("gate phases:", "  ".join(f"{lab} {acc[gate, i].mean():.0f}" for i, lab in enumerate(labels + ["gap"])))
        print("res  phases:", "  ".join(f"{lab} {acc[res, i].mean():.0f}" for i, lab in enumerate(labels + ["gap"])))
        print("step total (mean):", np.mean([tr[st + 1, 0, 0] - tr[st, 0, 0] for st in range(1, 7)]))
        return
    for st in range(1, 4):
        print(f"--- step {args.step + st}: total {tr[st + 1, 0, 0] - tr[st, 0, 0]} cycles")
        for ph in range(nphase):
            a, b, c, d = tr[st, ph]
            nxt = tr[st, ph + 1, 0] if ph + 1 < nphase else tr[st + 1, 0, 0]
            if ph < nphase - 1:
                print(f"{names[ph]} wait {b - a:6d}  mma {c - b:6d}  epi {d - c:6d}  gap {nxt - d:5d}")
            else:
                print(f"{names[ph]} total {d - a:6d} gap {nxt - d:5d}")
    avg = np.zeros((nphase, 3))
    for st in range(1, 7):
        for ph in range(nphase - 1):
            a, b, c, d = tr[st, ph]
            avg[ph] += [b - a, c - b, d - c]
    avg /= 6
    print("mean over 6 steps (cycles): wait %.0f mma %.0f epi %.0f per phase; gate wait %.0f res wait %.0f" % (
        avg[:-1, 0].mean(), avg[:-1, 1].mean(), avg[:-1, 2].mean(), avg[0:2 * L:2, 0].mean(), avg[1:2 * L:2, 0].mean()))


if __name__ == "__main__":
    main()
