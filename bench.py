#!/usr/bin/env python3
"""bench.py -- headline benchmark of the QPNet hot path on B200 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--scaling weak|strong]

Workload (config.workload): batch fast generation of synthetic 5 s utterances (1000 frames -> 109 999 samples each) with
the SI default model, random-init weights, synthetic WORLD-style aux features and F0 contours, sampling mode, 256
utterances per GPU -- the batch of BASELINE.json configs[3] on one GPU, the largest single-GPU configuration in
`configs`, and one launch of the two-group tcgen05 generator.  A "step" is one full pass of the hot path over that batch.
With N > 1 ranks (torchrun) every rank generates its own 256 utterances (utterance sharding, no collective: weak
scaling); `--scaling strong` shards configs[3]'s 256 utterances over the ranks instead (256 / N per GPU).  BASELINE configs[1] (32 utterances on one GPU) is measured in the same run and reported under
`configs1_32_utterances`, configs[2] (F0 x0.5 / x1.5) under `configs2_f0_scaled`, configs[0] (the CPU case) under
`configs0_cpu`, configs[4] (training) under `train`.

Metric: generated samples per second, whole job.
  value : inputs resident in HBM, device-timed with CUDA events (max over ranks).
  e2e   : the same through QPNet.batch_fast_generate with HOST buffers (pinned h / d in, int64 symbol arrays out),
          host<->device copies inside the timed region.
  roofline     : the persistent generator kernel against the measured HBM bandwidth.
  cpu_baseline : the CPU oracle port of the reference algorithm on the host cores (rank 0, N = 1 only), on a bounded
                 sample of the same workload.
  reference_on_b200 : the same port with CUDA tensors (eager PyTorch, the way the reference itself runs on a GPU).

--impl reference times the reference's CPU implementation of the path.  The reference is pure Python and cannot travel
to the GPU box (and must not be copied), so the arm runs the oracle port (oracle/qpnet_oracle.py, pinned to the
reference by tests/golden/) with every host thread.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UTTS_PER_GPU = 256       # configs[3]'s batch on one GPU; one launch of the two-group tcgen05 generator
UTTS_STRONG = 256        # configs[3]: 256 utterances over the box
FRAMES = 1000            # 5 s at 5 ms shift
FS = 22050
ALG_WEIGHT_BYTES = 47.25e6   # live weights in bf16, streamed once per step
ALG_STATE_BYTES = 31.7e3     # per utterance per step (FIFO taps, aux, offsets, symbol)
METRIC = "generated samples/sec (whole box)"


def build_inputs(n_utts, first_utt, frames, f0_factor=1.0):
    from qpnet_b200 import synth
    h = np.zeros((n_utts, synth.N_AUX, frames), np.float32)
    f0 = np.zeros((n_utts, frames), np.float64)
    n_list = []
    for b in range(n_utts):
        hs, f, n = synth.utterance(frames, first_utt + b, f0_factor)
        h[b] = hs.T
        f0[b] = f
        n_list.append(n)
    return h, f0, n_list


def host_dilated(f0):
    """caller-side d exactly as qpnet_decode.py:174-175 builds it (numpy fp64)."""
    d = np.ones(f0.shape) * FS
    d /= f0
    d /= 8
    return np.repeat(d, 110, axis=1)


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(prefix="qp_clocks_", suffix=".csv")
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nme, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
# the oracle port as a BASELINE (never as the product): CPU, and the same code with CUDA tensors
# ---------------------------------------------------------------------------------------------------------------
def oracle_sample(n_utts, steps, frames, threads, device="cpu"):
    """Time `steps` sample steps of the oracle port on `n_utts` utterances of the workload.  Returns (samples/s, s).
    The timed region starts after the port's priming (its constant-signal pass) and a 2-step warm-up call."""
    import torch
    from oracle import qpnet_oracle as orc
    torch.set_num_threads(threads)
    a = orc.Arch()
    p = orc.init_params(a, 0)
    if device != "cpu":
        p = {k: v.to(device) for k, v in p.items()}
    h, f0, n_list = build_inputs(n_utts, 0, frames)
    d = host_dilated(f0)
    g = torch.Generator().manual_seed(100)
    uni = torch.rand((n_utts, steps), generator=g)
    x = torch.full((n_utts, 1), a.Q // 2, dtype=torch.long)
    with torch.no_grad():
        orc.generate(a, p, x, torch.from_numpy(h), list(n_list), d, mode="sampling", uniforms=uni, max_steps=2)
        if device != "cpu":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        orc.generate(a, p, x, torch.from_numpy(h), list(n_list), d, mode="sampling", uniforms=uni, max_steps=steps)
        if device != "cpu":
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    return n_utts * steps / dt, dt


def config0_cpu(threads):
    """BASELINE configs[0] exactly: 1 s synthetic utterance (200 frames -> 21 999 samples), CPU,
    fp32, every host thread: (i) teacher-forced forward on one segment cut by the reference's segment rule, median of 3;
    (ii) batch-1 fast generation, sampling mode, 2 000 steady-state steps after priming."""
    import torch
    from oracle import qpnet_oracle as orc
    from qpnet_b200 import synth
    from qpnet_b200.train import segment_geometry
    torch.set_num_threads(threads)
    a = orc.Arch()
    p = orc.init_params(a, 0)
    frames = 200
    hs, f0, n = synth.utterance(frames, 0)
    d = host_dilated(f0[None])[0]
    R, bl, h_bs, x_bs = segment_geometry(float(d.max()), 20000, a.U, 1, a.rfF, a.rfA)
    x = torch.from_numpy(np.random.RandomState(0).randint(0, a.Q, size=(1, x_bs - 1))).long()
    h = torch.from_numpy(hs[:h_bs].T.copy())[None]
    dd = torch.from_numpy(d[: x_bs - 1].astype(np.float32))[None]
    times = []
    with torch.no_grad():
        for _ in range(3):
            t0 = time.perf_counter()
            orc.forward(a, p, x, h, dd, bl)
            times.append(time.perf_counter() - t0)
    fwd = float(np.median(times))
    v, dt = oracle_sample(1, 2000, frames, threads)
    return {"workload": "BASELINE configs[0]: 1 s synthetic utterance (200 frames, 21 999 samples), SI default model, CPU fp32",
            "cores": threads, "kind": "port",
            "forward": {"segment_bl": bl, "segment_samples": x_bs - 1, "seconds_median_of_3": fwd, "samples_per_s": bl / fwd,
                        "real_time_factor": bl / fwd / FS},
            "generate_b1": {"steps": 2000, "seconds": dt, "samples_per_s": v, "real_time_factor": v / FS}}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_per_step(kernel):
    """dram bytes per generator step of `kernel` from the committed ncu --set full captures (or None)."""
    path = os.path.join(ROOT, "profiles", "gen_kernel_traffic.json")
    if os.path.exists(path):
        try:
            j = json.load(open(path))
            return float(j["kernels"][kernel]["dram_bytes_per_step"])
        except Exception:
            return None
    return None


def kernel_of(n_utts):
    """The generator kernel a launch of n_utts utterances of the SI default model runs on (qp_generate.cu: wanted_kernel)."""
    if n_utts <= 32:
        return "qp::f2::f2_gen_kernel"
    return "qp::f3::f3_gen_kernel" if n_utts <= 128 else "qp::f3x2::f3x2_gen_kernel"


def roofline_of(kernel_s, n_utts, max_n, kernel, prime_steps):
    peak, peak_src = peaks()
    steps_per_launch = max_n + prime_steps
    alg = ALG_WEIGHT_BYTES + n_utts * ALG_STATE_BYTES
    achieved = steps_per_launch * alg / kernel_s / 1e9
    tps = ncu_traffic_per_step(kernel)
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": (tps * steps_per_launch if tps is not None else None), "peak_source": peak_src,
            "kernel": kernel, "us_per_sample_step": kernel_s / steps_per_launch * 1e6,
            "algorithmic_bytes_per_step": alg, "sample_steps_per_launch": steps_per_launch}


KERNEL_NOTE = ("SURVEY.md 8(d) models the step as weight-bandwidth bound (47.25 MB of bf16 weights per step + 31.7 KB of state "
               "per utterance); the step is in fact a chain of L + 3 = 19 dependent cross-SM exchanges (16 blocks, skip, 2 head "
               "layers) + sampling, each ~4 us per group of 128 utterances (profiles/r02h_*, r02x_*), so frac is small by construction")


def workload_config(n_gpus, per_gpu, scaling, frames):
    tot = per_gpu * n_gpus
    return {"workload": f"QPNet SI default, batch_fast_generate of {per_gpu} synthetic 5 s utterances per GPU ({frames} frames, "
                        f"{frames * 110 - 1} samples each), mode=sampling, extra_memory=False: "
                        + ("BASELINE configs[3] (256 utterances over the box) sharded over the ranks" if scaling == "strong" else
                           "the batch of BASELINE configs[3] on every GPU (the largest single-GPU configuration; "
                           "configs[1], 32 utterances, is reported under configs1_32_utterances)"),
            "utterances_per_gpu": per_gpu, "utterances_total": tot,
            "samples_per_utterance": frames * 110 - 1, "parallelism": f"utterance-sharded x{n_gpus}, no collective",
            "l2": "inputs are re-read per step and the generator's working set is re-packed per call; "
                  "a 512 MiB buffer is written between timed iterations to flush L2"}


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    per_gpu = UTTS_STRONG // args.gpus if args.scaling == "strong" else UTTS_PER_GPU
    sample_steps = args.ref_sample_steps
    vals = []
    for i in range(args.warmup + args.steps):
        v, dt = oracle_sample(per_gpu, sample_steps, FRAMES, threads)
        if i >= args.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([dt for _, dt in vals]) * 1e3)
    sample = (f"{sample_steps} steady-state sample steps of one rank's {per_gpu}-utterance batch per bench step (a rate over a "
              f"sample: the workload has 109 999 steps per utterance; the port's constant-signal priming is outside the timed "
              f"region); oracle port of qpnet.py:314-559, torch CPU fp32, ONE host")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "real_time_factor": value / FS,
            "config": workload_config(args.gpus, per_gpu, args.scaling, FRAMES),
            "cpu_baseline": {"value": value, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: 256 utterances per GPU; strong: BASELINE configs[3], 256 utterances over the ranks")
    ap.add_argument("--frames", type=int, default=FRAMES, help="(debug) frames per utterance")
    ap.add_argument("--utts", type=int, default=0, help="(debug) utterances per GPU")
    ap.add_argument("--ref-sample-steps", type=int, default=40)
    ap.add_argument("--cpu-sample-steps", type=int, default=240)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip configs[0] / [1] / [2], the eager-GPU port and the sample-match probe")
    ap.add_argument("--no-train", action="store_true", help="skip the train seg/s probe")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a B200; there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"

    import __graft_entry__
    if rank == 0:
        __graft_entry__.build()
    if world > 1:
        dist.barrier()
    from qpnet_b200 import _lib, ops
    from qpnet_b200.qpnet import QPNet, initialize

    torch.manual_seed(0)                      # identical random-init weights on every rank
    model = QPNet()
    model.apply(initialize)
    model = model.to(dev)
    model.philox_seed = 100
    frames = args.frames
    n_utts = args.utts or (UTTS_STRONG // world if args.scaling == "strong" else UTTS_PER_GPU)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def resident_inputs(n, first, f0_factor=1.0):
        h_np, f0_np, n_list = build_inputs(n, first, frames, f0_factor)
        h_dev = torch.from_numpy(h_np).to(dev)
        d64_dev, _ = ops.f0_to_dilated(torch.from_numpy(f0_np).to(dev), FS, 8, 110, want_f32=False)
        seed_dev = torch.full((n,), 128, dtype=torch.int64, device=dev)
        n_dev = torch.tensor(n_list, dtype=torch.int32, device=dev)
        ids = torch.arange(first, first + n, dtype=torch.int32, device=dev)      # corpus-level utterance ids key the Philox stream
        return (h_np, f0_np, n_list), (seed_dev, h_dev, d64_dev, n_dev, ids)

    def timed_resident(dev_in, n_list, steps, warmup, sampler=False):
        seed_dev, h_dev, d64_dev, n_dev, ids = dev_in
        max_n = max(n_list)

        def one():
            out, _ = model.generate_device(seed_dev, h_dev, d64_dev, n_dev, max_n, _lib.QP_MODE_SAMPLING, check_status=False,
                                           n_host=n_list, utt_ids=ids)
            return out
        for _ in range(warmup):
            one()
            flush.fill_(1)
        barrier()
        smp = ClockSampler(local) if (sampler and rank == 0) else None
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        launches = 0
        t0 = time.perf_counter()
        for i in range(steps):
            ev[i][0].record()
            out = one()
            ev[i][1].record()
            launches += model.last_launches
            flush.fill_(i)                      # L2 flush between timed iterations (outside the events)
        barrier()
        wall = time.perf_counter() - t0
        clocks = smp.stop() if smp else None
        status = _lib.lib.qp_workspace_status(model._last_ws.data_ptr(), torch.cuda.current_stream().cuda_stream)
        assert status == 0, _lib.lib.qp_last_error()
        ms = [a.elapsed_time(b) for a, b in ev]
        sym = out[:, :max_n].cpu().numpy()
        assert sym.min() >= 0 and sym.max() < 256 and len(np.unique(sym[:, :2000])) > 16, "degenerate output"
        return ms, launches, wall, clocks

    # ---------------- device-resident arm (value) ----------------
    (h_np, f0_np, n_list), dev_in = resident_inputs(n_utts, rank * n_utts)
    max_n = max(n_list)
    total_samples_rank = int(sum(n_list))
    dev_ms, launches, t_wall, clocks = timed_resident(dev_in, n_list, args.steps, args.warmup, sampler=True)
    t_dev = torch.tensor([sum(dev_ms) / 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    t_dev = float(t_dev.item())
    value = total_samples_rank * world * args.steps / t_dev

    # ---------------- end-to-end arm (public API, host buffers) ----------------
    h_pin = torch.from_numpy(h_np).pin_memory()
    d_pin = torch.from_numpy(host_dilated(f0_np)).pin_memory()
    x_seed = torch.full((n_utts, 1), 128, dtype=torch.long)
    h2d = h_pin.numel() * 4 + d_pin.numel() * 8 + n_utts * 8 + n_utts * 4
    d2h = n_utts * max_n * 4 + 4
    model.batch_fast_generate(x_seed, h_pin, list(n_list), d_pin)          # warm-up
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = model.batch_fast_generate(x_seed, h_pin, list(n_list), d_pin)
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = total_samples_rank * world * args.steps / float(t_e2e.item())
    assert len(res) == n_utts and all(len(r) == n for r, n in zip(res, sorted(n_list)))

    # ---------------- the other BASELINE configs (rank 0 of a 1-GPU run) ----------------
    extras = {}
    if world == 1 and not args.no_extras:
        # configs[1]: 32 utterances on one GPU (mma.sync cluster kernel)
        (_, _, nl32), in32 = resident_inputs(32, 0)
        ms32, _, _, _ = timed_resident(in32, nl32, 2, 1)
        s32 = sum(ms32) / len(ms32) / 1e3
        extras["configs1_32_utterances"] = {
            "workload": "BASELINE configs[1]: batch fast generation of 32 synthetic 5 s utterances on 1 x B200",
            "value": sum(nl32) / s32, "unit": "samples/s", "real_time_factor": sum(nl32) / s32 / FS, "steps": 2, "warmup": 1,
            "roofline": roofline_of(s32, 32, max(nl32), kernel_of(32), 16)}
        # configs[2]: the F0 contour scaled x0.5 (longest look-backs, ring depth 8 * ceil(max d)) and x1.5 (shortest)
        c2 = {}
        for fac in (0.5, 1.5):
            (_, f0f, nlf), inf = resident_inputs(n_utts, 0, fac)
            msf, _, _, _ = timed_resident(inf, nlf, 1, 1)
            sf = msf[0] / 1e3
            c2[f"f0_x{fac}"] = {"value": sum(nlf) / sf, "unit": "samples/s", "real_time_factor": sum(nlf) / sf / FS,
                                "utterances": n_utts, "max_dilated_factor": float(np.ceil((FS / f0f / 8).max())), "steps": 1, "warmup": 1,
                                "roofline": roofline_of(sf / ((n_utts + 255) // 256), min(n_utts, 256), max(nlf), kernel_of(min(n_utts, 256)), 16)}
        extras["configs2_f0_scaled"] = c2
        # free-running generation under shared pre-drawn uniforms against the CPU oracle: sample-match rate
        try:
            extras["sample_match"] = sample_match_probe(model, dev)
        except Exception as e:                      # a probe, never a reason to lose the benchmark line
            extras["sample_match"] = {"error": repr(e)}

    # ---------------- the metric's second half: train seg/s (BASELINE configs[4]), data-parallel steps ----------------
    train = None
    if not args.no_train:
        del flush
        torch.cuda.empty_cache()
        import importlib.util
        spec = importlib.util.spec_from_file_location("bench_train", os.path.join(ROOT, "tools", "bench_train.py"))
        bench_train = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(bench_train)
        train = bench_train.measure(dev, rank, world, steps=20, warmup=3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    kernel_s = (sum(dev_ms) / len(dev_ms)) / 1e3       # events bracket pack + generator; the generator is > 99.9 %
    n_launch = (n_utts + 255) // 256
    kernel = kernel_of(min(n_utts, 256))
    roofline = roofline_of(kernel_s / n_launch, min(n_utts, 256), max_n, kernel, 16)
    roofline["note"] = KERNEL_NOTE

    cpu = None
    if args.gpus == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, dt = oracle_sample(n_utts, args.cpu_sample_steps, frames, threads)
        cpu = {"value": v, "unit": "samples/s", "cores": threads, "kind": "port",
               "sample": f"{args.cpu_sample_steps} steady-state sample steps of the same {n_utts}-utterance batch "
                         f"({dt:.1f} s of CPU work), oracle port of qpnet.py:314-559, torch CPU fp32"}
        if not args.no_extras:
            try:
                # the reference's own way of running on a GPU: eager PyTorch, one small kernel per op (qpnet.py:446-516)
                vg, dtg = oracle_sample(32, 300, frames, threads, device=str(dev))
                extras["reference_on_b200"] = {
                    "value": vg, "unit": "samples/s", "real_time_factor": vg / FS, "kind": "port, eager PyTorch with CUDA tensors (fp32)",
                    "sample": f"300 steady-state sample steps of a 32-utterance batch ({dtg:.1f} s); BASELINE.md section 3 asks for this "
                              f"row (the reference with extra_memory=True on the same GPU); the unmodified reference cannot travel to the "
                              f"box, so its restatement runs the same per-sample loop of small ATen kernels"}
                extras["reference_on_b200"]["train_step_fp32"] = eager_train_step(dev)
            except Exception as e:
                extras["reference_on_b200"] = {"error": repr(e)}
            extras["configs0_cpu"] = config0_cpu(threads)

    line = {"metric": METRIC, "value": value, "unit": "samples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "real_time_factor": value / FS, "real_time_factor_per_gpu": value / FS / world,
            "config": workload_config(args.gpus, n_utts, args.scaling, frames),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "real_time_factor": e2e_value / FS},
            "gpu_launches": launches,
            "roofline": roofline, "cpu_baseline": cpu, "wall_s_timed_region": t_wall,
            "train": train}
    line.update(extras)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def sample_match_probe(model, dev):
    """Free-running generation under shared pre-drawn uniforms (north star): 4 utterances x 219 samples, this library (on
    the kernel the batch size selects and on the tcgen05 kernel) against the CPU oracle's own free run."""
    import torch
    from oracle import qpnet_oracle as orc
    a = orc.Arch()
    p = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    B, frames = 4, 2
    h, f0, n_list = build_inputs(B, 9000, frames)
    d = host_dilated(f0)
    n = frames * 110 - 1
    uni = torch.rand((B, n), generator=torch.Generator().manual_seed(100))
    x = torch.full((B, 1), a.Q // 2, dtype=torch.long)
    with torch.no_grad():
        ref = orc.generate(a, p, x, torch.from_numpy(h), [n] * B, d, mode="sampling", uniforms=uni)
    out = {}
    for kernel in ("fold2", "f3"):
        os.environ["QPNET_GEN_KERNEL"] = kernel
        res = model.batch_fast_generate(x, torch.from_numpy(h), [n] * B, d, None, "sampling", False, uniforms=uni)
        rates = [float((r == q).mean()) for r, q in zip(res, ref)]
        firsts = [int(np.nonzero(r != q)[0][0]) if (r != q).any() else n for r, q in zip(res, ref)]
        out[kernel] = {"match_rate": float(np.mean(rates)), "first_divergence": firsts, "samples_per_utterance": n}
    os.environ.pop("QPNET_GEN_KERNEL", None)
    out["note"] = ("random-init weights give near-uniform posteriors: a bf16 logit difference of ~0.02 flips an inverse-CDF draw "
                   "every ~10-40 steps and the trajectories then diverge; tests/test_gpu_parity.py asserts the per-step agreement "
                   "under the reference's history (> 0.95)")
    return out


def eager_train_step(dev):
    """One fp32 training step (forward + CE + backward + Adam) of the oracle port with CUDA tensors, B = 1, on the bench's
    training segment: the like-for-like bar for `train`."""
    import torch
    from oracle import qpnet_oracle as orc
    from qpnet_b200 import synth
    from qpnet_b200.train import segment_geometry
    a = orc.Arch()
    p = {k: v.to(dev).requires_grad_(True) for k, v in orc.init_params(a, 0).items()}
    frames = 260
    hs, f0, _ = synth.utterance(frames, 700)
    d = host_dilated(f0[None])[0].astype(np.float32)
    R, bl, h_bs, x_bs = segment_geometry(float(d.max()), 20000, a.U, 1, a.rfF, a.rfA)
    xq = torch.from_numpy(np.random.RandomState(0).randint(0, a.Q, size=(1, x_bs))).long()
    x, t = xq[:, :-1].to(dev), xq[:, 1:][:, -bl:].to(dev)
    h = torch.from_numpy(hs[:h_bs].T.copy())[None].to(dev)
    dd = torch.from_numpy(d[: x_bs - 1])[None]
    opt = torch.optim.Adam(list(p.values()), lr=1e-4)
    times = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        logits = orc.forward(a, p, x, h, dd, bl)
        loss = torch.nn.functional.cross_entropy(logits.reshape(-1, a.Q), t.reshape(-1))
        loss.backward()
        opt.step()
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    s = float(np.median(times[1:]))
    return {"ms_per_step": s * 1e3, "segments_per_s": 1.0 / s, "bl": bl, "kind": "port, eager PyTorch fp32 (TF32 off), autograd + torch Adam"}


if __name__ == "__main__":
    main()
