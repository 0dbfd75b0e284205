"""GPU parity at the BASELINE.json sizes (-m gpu).

tests/test_gpu_parity.py pins the kernels on shapes the oracle finishes in a second; this file repeats the comparisons
at the sizes bench.py measures, through the same public class API (ctypes -> C ABI):

* teacher-forced logits AND gradients of the full-width model on the bench's training segment (bl ~ 19 939), fp32 SIMT
  path and bf16 tcgen05 path, against ``oracle.forward`` + autograd (qpnet.py:239-312, qpnet_train.py:526-529);
* the benchmarked generator kernels with a full group of utterances, F0 x0.5 (M ~ 123, ring depth 1024) for more steps
  than the deepest ring holds, per-step logits under forced symbols against ``oracle.generate`` (qpnet.py:446-516);
* the Adam update of the bf16 tensor-core training path (qpnet_train.py:526-531);
* free-running generation under shared pre-drawn uniforms: sample-match rate and first divergence per kernel.
"""
import os

import numpy as np
import pytest
import torch

from oracle import qpnet_oracle as orc
from qpnet_b200 import ops, synth
from qpnet_b200.qpnet import QPNet
from qpnet_b200.train import Trainer, segment_geometry
from tests import cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "the gpu suite needs a B200"
    return torch.device("cuda:0")


def _model(kw, p, dev, tensor_cores=False):
    m = QPNet(**kw)
    m.load_state_dict(p)
    m = m.to(dev)
    m.tensor_cores = tensor_cores
    return m


def bench_segment(a, seed, batch_length=20000):
    """The segment tools/bench_train.py trains on, cut the reference's way (qpnet_train.py:268-303)."""
    frames = 260
    hs, f0, _ = synth.utterance(frames, 700 + seed)
    d = cases.d_from_f0(f0).astype(np.float32)
    R, bl, h_bs, x_bs = segment_geometry(float(d.max()), batch_length, a.U, 1, a.rfF, a.rfA)
    wav = synth.noise_waveform(x_bs, seed)
    xq = torch.from_numpy(orc.encode_mu_law(wav.astype(np.float64)))
    x, t = xq[None, :-1].contiguous(), xq[None, 1:].contiguous()
    h = torch.from_numpy(hs[:h_bs].T.copy())[None]
    dd = torch.from_numpy(d[: x_bs - 1].copy())[None]
    return x, h, dd, t[:, -bl:].contiguous(), bl, R


@pytest.fixture(scope="module")
def segment_oracle():
    """Oracle logits and gradients of the bench segment (about a minute of CPU work, shared by both paths)."""
    torch.set_num_threads(os.cpu_count() or 1)
    a = orc.Arch()
    p = orc.init_params(a, 12, 0.05)
    x, h, d, t, bl, R = bench_segment(a, 0)
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    want = orc.forward(a, pr, x, h, d, bl)
    loss = torch.nn.functional.cross_entropy(want.reshape(-1, a.Q), t.reshape(-1))
    loss.backward()
    grads = {k: (v.grad.clone() if v.grad is not None else None) for k, v in pr.items()}
    return a, p, x, h, d, t, bl, R, want.detach(), float(loss), grads


@pytest.mark.parametrize("tensor_cores", [False, True])
def test_forward_backward_at_bench_segment_vs_oracle(dev, segment_oracle, tensor_cores):
    """Full-width teacher-forced pass on the segment the training benchmark uses (bl = 19 939, 20 240 samples).
    fp32 SIMT path: logits 2e-4 absolute; gradients are sums of ~20 000 random-sign rows, so fp32 summation order alone
    moves them by ~4e-4 relative L2 (tools/diag_scale_grads.py, DIAG_F64=1 shows the oracle's own fp32-vs-fp64 distance):
    bar 2e-3 relative L2 and 6e-3 of the tensor's max (measured: 1.3e-3 / 4.1e-3 worst, causal.conv.weight).
    bf16 tcgen05 path: logits 0.05 absolute (the bar of the small shapes; measured 0.031); gradients: relative L2 per
    tensor against the fp32 oracle below 0.10 (measured: median 0.026, worst 0.080 on causal.conv.weight, 0.05 on the
    sigmoid-side gate / aux matrices: bf16 rounding of dgate, dX and the saved activations)."""
    a, p, x, h, d, t, bl, R, want, want_loss, grads = segment_oracle
    assert bl > 19000 and x.shape[1] == R + bl
    m = _model({}, p, dev, tensor_cores=tensor_cores)
    got = m(x.to(dev), h.to(dev), d.to(dev), torch.tensor([bl], device=dev))
    assert got.shape == (1, bl, a.Q)
    err = float((got.detach().cpu() - want).abs().max())
    loss, dl = ops.cross_entropy(got.detach(), t.to(dev))
    got.backward(dl)
    print(f"bench segment bl={bl} tensor_cores={tensor_cores}: max |dlogit| = {err:.3e}, loss {float(loss):.6f} vs {want_loss:.6f}")
    assert err < (0.05 if tensor_cores else 2e-4), err
    assert abs(float(loss) - want_loss) < (2e-3 if tensor_cores else 2e-5)
    last = f"resA_1x1.{len(a.dilA) - 1}"
    worst = ("", 0.0)
    for k, prm in m.named_parameters():
        ref = grads[k]
        if k.startswith(last):
            assert ref is None or float(ref.abs().max()) == 0.0          # dead projection (C7)
            continue
        g = prm.grad.detach().cpu()
        e = float((g - ref).norm() / ref.norm().clamp_min(1e-12))
        if e > worst[1]:
            worst = (k, e)
        if not tensor_cores:
            em = float((g - ref).abs().max() / ref.abs().max().clamp_min(1e-12))
            assert em < 6e-3, (k, em)
    print(f"  worst gradient error (relative L2): {worst[0]} {worst[1]:.3e}")
    assert worst[1] < (0.10 if tensor_cores else 2e-3), worst


def _forced_case(a, B, frames, fac, steps, seed0):
    h = np.zeros((B, a.A, frames), np.float32)
    d = np.zeros((B, frames * a.U), np.float64)
    for b in range(B):
        hs, f0, _ = synth.utterance(frames, seed0 + b, fac, a.A)
        h[b] = hs.T
        d[b] = cases.d_from_f0(f0)
    x = torch.full((B, 1), a.Q // 2, dtype=torch.long)
    forced = torch.from_numpy(np.random.RandomState(seed0).randint(0, a.Q, size=(B, steps))).long()
    return x, torch.from_numpy(h), d, forced


@pytest.fixture(scope="module")
def ringwrap_oracle():
    """32 utterances, F0 x0.5, 2 300 forced steps on the CPU oracle (deepest ring: 8 * M = 984 slots -> 1024)."""
    torch.set_num_threads(os.cpu_count() or 1)
    a = orc.Arch()
    p = orc.init_params(a, 41, 0.05)
    B, frames, steps = 32, 21, 2300
    x, h, d, forced = _forced_case(a, B, frames, 0.5, steps, 1200)
    lg = []
    with torch.no_grad():
        orc.generate(a, p, x, h, [steps] * B, d, mode="argmax", force=forced, logits_out=lg, max_steps=steps)
    return a, p, x, h, d, forced, steps, torch.stack(lg, dim=1)


@pytest.mark.parametrize("kernel", ["f3", "f3x2", "fold2"])
def test_generator_full_group_ring_wrap_vs_oracle(dev, monkeypatch, ringwrap_oracle, kernel):
    """The benchmarked kernels at the benchmark's batch (32 utterances), lowest pitch (F0 x0.5: look-backs up to
    8 * 123 samples), for 2 300 steps so that every past-tap ring wraps at least twice; per-step logits of every
    utterance under forced symbols against the oracle, 0.06 absolute."""
    a, p, x, h, d, forced, steps, want = ringwrap_oracle
    assert int(np.ceil(d.max())) >= 60          # ring depth 8 * 66 = 528 -> 1024 slots
    monkeypatch.setenv("QPNET_GEN_KERNEL", kernel)
    m = _model({}, p, dev)
    B = x.shape[0]
    res, got = m.batch_fast_generate(x, h, [steps] * B, d, None, "argmax", False, force=forced, return_logits=True)
    err = (got.cpu() - want).abs().amax(dim=(0, 2))          # per step
    print(f"{kernel}: B={B} steps={steps} max |dlogit| = {float(err.max()):.4f} (first 300 steps {float(err[:300].max()):.4f}, "
          f"last 300 {float(err[-300:].max()):.4f})")
    assert float(err.max()) < 0.06, (kernel, float(err.max()), int(err.argmax()))
    assert all(len(r) == steps for r in res)


@pytest.mark.parametrize("B,fac", [(128, 1.0), (77, 1.5), (256, 1.0), (150, 0.5)])
def test_generator_large_group_vs_oracle(dev, monkeypatch, B, fac):
    """The tcgen05 generators with a full 128-utterance group and a ragged one (qp_generate_f3.cu), two full groups and a
    full + ragged pair (qp_generate_f3x2.cu): per-step logits of EVERY utterance under forced symbols against the oracle,
    0.06 absolute; 260 steps read the first adaptive rings back."""
    torch.set_num_threads(os.cpu_count() or 1)
    a = orc.Arch()
    p = orc.init_params(a, 43, 0.05)
    frames, steps = 3, 260
    x, h, d, forced = _forced_case(a, B, frames, fac, steps, 1500)
    lg = []
    with torch.no_grad():
        orc.generate(a, p, x, h, [steps] * B, d, mode="argmax", force=forced, logits_out=lg, max_steps=steps)
    want = torch.stack(lg, dim=1)
    monkeypatch.setenv("QPNET_GEN_KERNEL", "f3")
    m = _model({}, p, dev)
    res, got = m.batch_fast_generate(x, h, [steps] * B, d, None, "argmax", False, force=forced, return_logits=True)
    err = (got.cpu() - want).abs().amax(dim=(1, 2))          # per utterance
    print(f"f3: B={B} f0 x{fac} max |dlogit| = {float(err.max()):.4f} (utterance {int(err.argmax())})")
    assert float(err.max()) < 0.06, err


@pytest.mark.parametrize("kernel", ["f3", "fold2", "generic"])
def test_generator_free_running_sample_match(dev, monkeypatch, kernel):
    """Free-running generation of the SI default model under shared pre-drawn uniforms (north star: 'reporting the
    sample-match rate'): 4 utterances x 1 100 samples against the oracle's own free run.  Trajectories are chaotic once
    a symbol flips, so the rate and the first divergence are reported; asserted: the sampler agrees with the oracle
    for as long as the histories agree (first divergence beyond the seed region) and stays in range."""
    torch.set_num_threads(os.cpu_count() or 1)
    a = orc.Arch()
    p = orc.init_params(a, 47, 0.05)
    B, frames = 4, 10
    n = frames * a.U - 1
    x, h, d, _ = _forced_case(a, B, frames, 1.0, 1, 1700)
    uni = torch.rand((B, n), generator=torch.Generator().manual_seed(100))
    with torch.no_grad():
        ref = orc.generate(a, p, x, h, [n] * B, d, mode="sampling", uniforms=uni)
    monkeypatch.setenv("QPNET_GEN_KERNEL", kernel)
    m = _model({}, p, dev)
    res = m.batch_fast_generate(x, h, [n] * B, d, None, "sampling", False, uniforms=uni)
    rates, firsts = [], []
    for b in range(B):
        neq = np.nonzero(res[b] != ref[b])[0]
        firsts.append(int(neq[0]) if len(neq) else n)
        rates.append(float((res[b] == ref[b]).mean()))
        assert res[b].min() >= 0 and res[b].max() < a.Q
    print(f"{kernel}: free-running sample match {np.mean(rates):.4f} per utterance {['%.3f' % r for r in rates]}, "
          f"first divergence {firsts} of {n}")
    assert min(firsts) >= 1


def test_training_step_bf16_tensor_cores_adam_update(dev):
    """One Trainer.step on the bf16 tcgen05 path (forward AND backward GEMMs on tensor cores) against the reference's
    recipe on the CPU oracle (CE on the last bl logits, autograd, Adam lr 1e-4, qpnet_train.py:526-531).  The first
    Adam step moves every element by -lr * g / (|g| + eps): compared where the oracle gradient is at least 5 % of its
    tensor's max (bf16 operand noise cannot flip those), 99.9 % of them within 2e-6 of the oracle's update."""
    kw = dict(n_resch=64, n_skipch=64)
    a = orc.Arch(**kw)
    p = orc.init_params(a, 31, 0.1)
    _, _, _, x, h, d, t, bl = cases.forward_inputs("small_s2_b1")
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    want = orc.forward(a, pr, x, h, d, bl)
    want_loss = torch.nn.functional.cross_entropy(want.reshape(-1, a.Q), t.reshape(-1))
    want_loss.backward()
    m = _model(kw, p, dev, tensor_cores=True)
    tr = Trainer(m, lr=1e-4)
    before = {k: v.detach().clone() for k, v in m.named_parameters()}
    loss = tr.step(x.to(dev), h.to(dev), d.to(dev), t.to(dev), bl)
    assert abs(float(loss) - float(want_loss)) < 5e-3
    last = f"resA_1x1.{len(a.dilA) - 1}"
    checked = good = 0
    for k, prm in m.named_parameters():
        ref = pr[k].grad
        if k.startswith(last) or ref is None or float(ref.abs().max()) == 0.0:
            continue
        ref = ref.to(dev)
        delta = prm.detach() - before[k]
        big = ref.abs() > 0.05 * ref.abs().max()
        upd = -1e-4 * ref / (ref.abs() + 1e-8)
        ok = (delta - upd)[big].abs() < 2e-6
        checked += int(big.sum())
        good += int(ok.sum())
    print(f"bf16 training step: {good} of {checked} significant elements moved like the oracle's Adam step")
    assert checked > 1000 and good >= 0.999 * checked
    l0 = float(loss)
    for _ in range(5):
        loss = tr.step(x.to(dev), h.to(dev), d.to(dev), t.to(dev), bl)
    assert float(loss) < l0


@pytest.mark.parametrize("kernel,B", [("f3", 40), ("f3x2", 140), ("fold2", 5), ("generic", 5)])
def test_generator_pcm16_output_stage(dev, monkeypatch, kernel, B):
    """The generator's PCM output stage (QpGenerateArgs.out_pcm; qpnet_decode.py:315-318: decode_mu_law * 32768, clipped,
    int16): equal, sample for sample, to the oracle's write-out of the symbols the same launch produced -- on the tcgen05
    kernel (in-kernel table) and on the kernels the C ABI post-processes."""
    a = orc.Arch()
    p = orc.init_params(a, 51, 0.05)
    frames = 1
    x, h, d, _ = _forced_case(a, B, frames, 1.0, 1, 1900)
    n = frames * a.U - 1
    monkeypatch.setenv("QPNET_GEN_KERNEL", kernel)
    m = _model({}, p, dev)
    m.philox_seed = 3
    pcm = torch.zeros((B, n), dtype=torch.int16, device=dev)
    seed = torch.full((B,), a.Q // 2, dtype=torch.int64, device=dev)
    sym, _ = m.generate_device(seed, h.to(dev), torch.from_numpy(d).to(dev), torch.full((B,), n, dtype=torch.int32, device=dev), n,
                               pcm_out=pcm)
    sym, pcm = sym.cpu().numpy(), pcm.cpu().numpy()
    assert sym.min() >= 0 and sym.max() < a.Q and len(np.unique(sym)) > 16
    for b in range(B):
        np.testing.assert_array_equal(pcm[b], orc.decode_pcm16(sym[b, :n].astype(np.int64)))


def test_deep_preset_full_width_on_tcgen05_generator(dev, monkeypatch):
    """`Rd10Rr3Ed4Er1` (param_model.py:65-71: 30 fixed blocks with dilations up to 512 + 4 adaptive ones) at FULL width
    (512 / 256 channels) on the tcgen05 generator (any depth up to QP_MAX_LAYERS; past-tap rings up to 1024 slots for the
    fixed blocks): per-step logits under forced symbols against the oracle; 620 steps read the 512-deep fixed rings back.
    Bar: 0.10 absolute -- the 0.06 of the 16-block model scaled by sqrt(34 / 16) is 0.0875 for per-block noise that adds
    like a random walk; the folded kernel also rounds the folded products G, H of every block to bf16, and the measured
    maximum over 3 x 620 x 256 logits sits at 0.083 - 0.090 depending on the rounding realisation."""
    torch.set_num_threads(os.cpu_count() or 1)
    kw = dict(dilationF_depth=10, dilationF_repeat=3, dilationA_depth=4, dilationA_repeat=1)
    a = orc.Arch(**kw)
    assert len(a.dilF) == 30 and max(a.dilF) == 512
    p = orc.init_params(a, 19, 0.05)
    B, frames, steps = 3, 6, 620
    x, h, d, forced = _forced_case(a, B, frames, 1.0, steps, 2100)
    lg = []
    with torch.no_grad():
        orc.generate(a, p, x, h, [steps] * B, d, mode="argmax", force=forced, logits_out=lg, max_steps=steps)
    want = torch.stack(lg, dim=1)
    monkeypatch.setenv("QPNET_GEN_KERNEL", "f3")
    m = _model(kw, p, dev)
    from qpnet_b200 import _lib
    res, got = m.batch_fast_generate(x, h, [steps] * B, d, None, "argmax", False, force=forced, return_logits=True)
    err = (got.cpu() - want).abs().amax(dim=(0, 2))
    print(f"deep preset, full width, f3: max |dlogit| = {float(err.max()):.4f} (steps 0-99 {float(err[:100].max()):.4f}, last 100 {float(err[-100:].max()):.4f})")
    assert float(err.max()) < 0.10, float(err.max())


@pytest.mark.parametrize("kernel", ["f3", "fold2"])
def test_generator_wide_aux_vs_oracle(dev, monkeypatch, kernel):
    """n_aux = 44 (more than the 40-float staging pitch the mma.sync kernel used in round 1, below its 48-channel limit):
    per-step logits under forced symbols against the oracle on both cluster kernels -- the aux rows of neighbouring
    utterances must not alias."""
    kw = dict(n_aux=44)
    a = orc.Arch(**kw)
    p = orc.init_params(a, 53, 0.05)
    B, frames, steps = 5, 2, 130
    x, h, d, forced = _forced_case(a, B, frames, 1.0, steps, 2300)
    lg = []
    with torch.no_grad():
        orc.generate(a, p, x, h, [steps] * B, d, mode="argmax", force=forced, logits_out=lg, max_steps=steps)
    want = torch.stack(lg, dim=1)
    monkeypatch.setenv("QPNET_GEN_KERNEL", kernel)
    m = _model(kw, p, dev)
    res, got = m.batch_fast_generate(x, h, [steps] * B, d, None, "argmax", False, force=forced, return_logits=True)
    err = float((got.cpu() - want).abs().max())
    print(f"n_aux 44, kernel {kernel}: max |dlogit| = {err:.4f}")
    assert err < 0.06, err
