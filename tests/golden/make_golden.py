#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the authoring container only (it needs /root/reference; the GPU box does not
have it):   python tests/golden/make_golden.py

The reference class is imported read-only from /root/reference/src/nets/qpnet.py and
driven through its public entry points (``forward``, ``batch_fast_generate``,
``_dilated_index``, ``_generate_dilated_index``, ``encode_mu_law``, ``decode_mu_law``).
Parameters come from ``oracle.qpnet_oracle.init_params`` (deterministic, seeded) and
are loaded with ``load_state_dict`` so that tests can rebuild the identical model
without the reference.  ``Categorical.sample`` is replaced, for the sampling fixtures
only, by an inverse-CDF draw on pre-drawn uniforms (the reference's global-RNG
multinomial cannot be shared with another implementation).
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src/nets")
warnings.filterwarnings("ignore")

import qpnet as ref  # noqa: E402  (the reference)
from oracle import qpnet_oracle as orc  # noqa: E402
from qpnet_b200 import synth  # noqa: E402

torch.set_num_threads(8)
torch.manual_seed(0)


def ref_model(arch_kw, params):
    m = ref.QPNet(**arch_kw)
    m.load_state_dict(params)
    m.eval()
    return m


def d_from_f0(f0, f32=False):
    d = orc.extend_time(orc.dilated_factor(f0, synth.FS, synth.DENSE_FACTOR), synth.UPSAMPLING)
    return d.astype(np.float32) if f32 else d


# ---------------------------------------------------------------- G1: indices
def golden_indices():
    tiny = ref.QPNet(n_resch=4, n_skipch=4)
    out = {}
    for factor in (1.0, 0.5, 1.5):
        for n in (1100, 20020, 30030):
            frames = n // synth.UPSAMPLING
            f0 = np.stack([synth.f0_contour(frames, u) * factor for u in (0, 1)])
            d64 = np.stack([d_from_f0(f) for f in f0])
            d32 = torch.from_numpy(d64).float()
            for dil in (1, 2, 4, 8):
                key = f"f{factor}_n{n}_d{dil}"
                a = tiny._dilated_index(d32, dil, 1, True)[:, 0].numpy()
                b = tiny._dilated_index(d64, dil, 1, False)[:, 0]
                c = tiny._generate_dilated_index(d32, dil, 1, True)[:, 0].numpy()
                e = tiny._generate_dilated_index(d64, dil, 1, False)[:, 0]
                pos = np.arange(-n, 0)
                out[key + "_tf32"] = (a - pos).astype(np.int16)   # store look-back, compresses well
                out[key + "_tf64"] = (b - pos).astype(np.int16)
                out[key + "_g32"] = c.astype(np.int16)
                out[key + "_g64"] = e.astype(np.int16)
    np.savez_compressed(os.path.join(HERE, "indices.npz"), **out)
    print("indices:", len(out), "arrays")


# ---------------------------------------------------------------- G2/G3: forward + grads
def tf_case(arch_kw, seed, bias_std, frames, bl, f0_factor=1.0, grads=False):
    a = orc.Arch(**arch_kw)
    p = orc.init_params(a, seed, bias_std)
    m = ref_model(arch_kw, p)
    hs, f0, _ = synth.utterance(frames, seed, f0_factor, a.A)
    T = frames * a.U
    d = torch.from_numpy(d_from_f0(f0)).float()[None, :T]
    rs = np.random.RandomState(seed)
    x = torch.from_numpy(rs.randint(0, a.Q, size=(1, T))).long()
    t = torch.from_numpy(rs.randint(0, a.Q, size=(1, bl))).long()
    h = torch.from_numpy(hs.T.copy())[None]
    logits = m(x, h, d, torch.tensor([bl]))
    res = {"logits": logits.detach().numpy()[0]}
    if grads:
        loss = torch.nn.functional.cross_entropy(logits.reshape(-1, a.Q), t.reshape(-1))
        loss.backward()
        res["loss"] = np.float32(loss.item())
        for name, prm in m.named_parameters():
            res["grad/" + name] = (prm.grad.numpy() if prm.grad is not None
                                   else np.zeros(0, np.float32))
        res["target"] = t.numpy()[0]
    res["x"] = x.numpy()[0].astype(np.int16)
    return res


SMALL = dict(n_resch=32, n_skipch=16)
FULL = dict()


def golden_forward():
    out = {}
    cases = [("small_s1_b0", SMALL, 1, 0.0, 12, 330, 1.0, True),
             ("small_s2_b1", SMALL, 2, 0.1, 14, 330, 1.0, True),
             ("small_s3_b1_f05", SMALL, 3, 0.1, 22, 220, 0.5, False),
             ("full_s4_b1", FULL, 4, 0.05, 12, 330, 1.0, False)]
    for name, kw, seed, bstd, frames, bl, fac, grads in cases:
        r = tf_case(kw, seed, bstd, frames, bl, fac, grads)
        for k, v in r.items():
            out[f"{name}/{k}"] = v
        out[f"{name}/meta"] = np.array([seed, frames, bl], np.int64)
        out[f"{name}/meta_f"] = np.array([bstd, fac], np.float64)
        print("forward", name, r["logits"].shape, float(np.abs(r["logits"]).max()))
    np.savez_compressed(os.path.join(HERE, "forward.npz"), **out)


# ---------------------------------------------------------------- G4: generation
class _ShimCategorical:
    """Inverse-CDF stand-in for torch.distributions.Categorical (sampling fixtures)."""
    uniforms = None
    step = 0
    rows = None

    def __init__(self, probs):
        self.probs = probs

    def sample(self):
        cls = _ShimCategorical
        u = cls.uniforms[cls.rows_alive(), cls.step]
        cdf = torch.cumsum(self.probs, dim=-1)
        k = torch.searchsorted(cdf, u.reshape(-1, 1).to(cdf.dtype), right=True).reshape(-1)
        cls.step += 1
        return torch.clamp(k, max=self.probs.shape[-1] - 1)

    @classmethod
    def rows_alive(cls):
        return cls.rows


def gen_case(arch_kw, seed, bias_std, frames_list, mode, extra_memory, f0_factor=1.0):
    a = orc.Arch(**arch_kw)
    p = orc.init_params(a, seed, bias_std)
    m = ref_model(arch_kw, p)
    B = len(frames_list)
    Fmax = max(frames_list)
    h = np.zeros((B, a.A, Fmax), np.float32)
    d = np.zeros((B, Fmax * a.U), np.float64)
    n_list = []
    for b, fr in enumerate(frames_list):
        hs, f0, n = synth.utterance(fr, 100 * seed + b, f0_factor, a.A)
        h[b, :, :fr] = hs.T
        d[b, : fr * a.U] = d_from_f0(f0)
        n_list.append(n)
    x = torch.full((B, 1), a.Q // 2, dtype=torch.long)
    g = torch.Generator().manual_seed(100)
    uniforms = torch.rand((B, max(n_list)), generator=g)
    # rows alive bookkeeping: the reference drops finished utterances (shortest first)
    order = np.argsort(np.array(n_list), kind="stable")
    _ShimCategorical.uniforms = uniforms
    _ShimCategorical.step = 0

    def rows_alive():
        step = _ShimCategorical.step
        return torch.tensor([b for b in range(B) if n_list_orig[b] > step])
    n_list_orig = list(n_list)
    _ShimCategorical.rows_alive = classmethod(lambda cls: rows_alive())
    saved = torch.distributions.Categorical
    torch.distributions.Categorical = _ShimCategorical
    try:
        with torch.no_grad():
            dd = torch.from_numpy(d).float() if extra_memory else d
            res = m.batch_fast_generate(x, torch.from_numpy(h), list(n_list), dd,
                                        None, mode, extra_memory)
    finally:
        torch.distributions.Categorical = saved
    # reference returns ascending-length (finish) order
    by_input = [None] * B
    for r, b in zip(res, order):
        assert len(r) == n_list_orig[b]
        by_input[b] = r
    return by_input, uniforms.numpy(), n_list_orig


def golden_generate():
    out = {}
    cases = [("small_argmax", SMALL, 5, 0.1, [3, 5, 4], "argmax", False, 1.0),
             ("small_sampling", SMALL, 6, 0.1, [4, 3, 5], "sampling", False, 1.0),
             ("small_sampling_xm", SMALL, 6, 0.1, [4, 3, 5], "sampling", True, 1.0),
             ("small_sampling_f05", SMALL, 7, 0.1, [6, 6], "sampling", False, 0.5),
             ("small_sampling_f15", SMALL, 8, 0.1, [5, 4], "sampling", False, 1.5),
             ("full_sampling", FULL, 9, 0.05, [2, 3], "sampling", False, 1.0)]
    for name, kw, seed, bstd, frames, mode, xm, fac in cases:
        res, uni, n_list = gen_case(kw, seed, bstd, frames, mode, xm, fac)
        for b, r in enumerate(res):
            out[f"{name}/sym{b}"] = r.astype(np.int16)
        out[f"{name}/uniforms"] = uni.astype(np.float32)
        out[f"{name}/meta"] = np.array([seed] + frames, np.int64)
        out[f"{name}/meta_f"] = np.array([bstd, fac], np.float64)
        print("generate", name, [len(r) for r in res], [int(r[:8].sum()) for r in res])
    np.savez_compressed(os.path.join(HERE, "generate.npz"), **out)


# ---------------------------------------------------------------- G5: mu-law
def golden_mulaw():
    xs = np.array([-1, -.5, -.01, 0, .01, .5, 1], np.float64)
    rs = np.random.RandomState(0)
    xr = np.clip(rs.randn(4096) * 0.3, -1, 1)
    np.savez_compressed(os.path.join(HERE, "mulaw.npz"),
                        x_known=xs, enc_known=ref.encode_mu_law(xs),
                        x_rand=xr, enc_rand=ref.encode_mu_law(xr),
                        dec_all=ref.decode_mu_law(np.arange(256)))
    print("mulaw enc known:", ref.encode_mu_law(xs))


# ---------------------------------------------------------------- G7: decode front / back end
def _reference_functions(path, names, namespace):
    """Compile the named top-level functions of a reference script WITHOUT importing the script (its module-level
    imports need h5py / docopt, which are not installed): the function bodies are the reference's, unmodified."""
    import ast
    tree = ast.parse(open(path).read())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            code = compile(ast.Module(body=[node], type_ignores=[]), path, "exec")
            exec(code, namespace)
    missing = [n for n in names if n not in namespace]
    assert not missing, missing
    return namespace


def golden_decode():
    """decode_generator's per-utterance arithmetic (qpnet_decode.py:158-200) and the write-out (315-318), driven with
    the reference's own pad_list / _dilated_factor / _batch_f0 / extend_time / decode_mu_law and sklearn's
    StandardScaler (268-269) on seeded synthetic WORLD-style features."""
    import copy
    from numpy.matlib import repmat
    from sklearn.preprocessing import StandardScaler
    ns = {"np": np, "copy": copy, "repmat": repmat}
    _reference_functions("/root/reference/src/bin/qpnet_decode.py", ["pad_list", "_dilated_factor", "_batch_f0"], ns)
    _reference_functions("/root/reference/src/utils/utils.py", ["extend_time"], ns)
    out = {}
    rs = np.random.RandomState(5)
    D = 39
    mean = rs.randn(D) * 2.0
    scale = rs.rand(D) * 3.0 + 0.5
    mean[0], scale[0] = 0.0, 1.0                     # calc_stats.py:29-33 leaves the U/V flag unscaled
    scaler = StandardScaler()
    scaler.mean_, scaler.scale_, scaler.n_features_in_ = mean, scale, D
    out["mean"], out["scale"] = mean, scale
    for factor in (1.0, 0.5, 1.5):
        feats = []
        for u, frames in enumerate((7, 4, 9)):
            hs, f0, _ = synth.utterance(frames, 300 + u, 1.0, D)
            raw = rs.randn(frames, D)
            raw[:, 0] = (rs.rand(frames) > 0.3)
            raw[:, 1] = f0
            if u == 1:
                raw[2, 1] = 0.0                      # an unvoiced hole: f0 == 0 -> fs / dense_factor (101)
            feats.append(raw)
        batch_h, batch_d = [], []
        for raw in feats:                            # the loop body of decode_generator, 163-183
            h = raw.copy()
            h[:, 1] = h[:, 1] * factor
            d = ns["_dilated_factor"](ns["_batch_f0"](h), synth.FS, synth.DENSE_FACTOR)
            d = ns["extend_time"](np.expand_dims(d, -1), synth.UPSAMPLING)
            h = scaler.transform(h)
            batch_h += [h]
            batch_d += [d]
        bh = torch.from_numpy(ns["pad_list"](batch_h)).float().transpose(1, 2).numpy()   # 187, 190
        bd = ns["pad_list"](batch_d).squeeze(-1)                                          # 188, 194
        for u, raw in enumerate(feats):
            out[f"f{factor}/raw{u}"] = raw
        out[f"f{factor}/h"] = bh
        out[f"f{factor}/d"] = bd
    sym = np.arange(256, dtype=np.int64)
    wav = ref.decode_mu_law(sym, 256)
    out["pcm_all_symbols"] = np.clip(wav * 32768, -32768, 32767).astype(np.int16)        # 317-319
    np.savez_compressed(os.path.join(HERE, "decode.npz"), **out)
    print("decode.npz:", len(out), "arrays")


# ---------------------------------------------------------------- G8: checkpoint written the reference's way
def golden_checkpoint():
    """A small reference model + Adam state after one step, saved exactly like qpnet_train.py:336-352 / 389."""
    import argparse
    arch = dict(n_resch=8, n_skipch=8)
    a = orc.Arch(**arch)
    m = ref_model(arch, orc.init_params(a, 12, 0.1))
    m.train()
    opt = torch.optim.Adam(m.parameters(), lr=1e-4)
    T = 6 * a.U
    hs, f0, _ = synth.utterance(6, 77, 1.0, a.A)
    x = torch.from_numpy(np.random.RandomState(3).randint(0, a.Q, size=(1, T))).long()
    h = torch.from_numpy(hs.T.copy())[None]
    d = torch.from_numpy(d_from_f0(f0)).float()[None, :T]
    out = m(x, h, d, torch.tensor([100]))
    torch.nn.functional.cross_entropy(out.reshape(-1, a.Q), x[0, -100:]).backward()
    opt.step()
    checkpoint = {"model": m.state_dict(), "optimizer": opt.state_dict(), "iterations": 7}
    torch.save(checkpoint, os.path.join(HERE, "checkpoint-7.pkl"))
    ns = argparse.Namespace(n_quantize=256, n_aux=39, n_resch=8, n_skipch=8, dilationF_depth=4, dilationF_repeat=3,
                            dilationA_depth=4, dilationA_repeat=1, kernel_size=2, upsampling_factor=110, lr=1e-4, seed=1)
    torch.save(ns, os.path.join(HERE, "model.conf"))
    print("checkpoint-7.pkl", os.path.getsize(os.path.join(HERE, "checkpoint-7.pkl")), "bytes")


# ---------------------------------------------------------------- G9: the training segmenter, run by the reference
def segmenter_utterances():
    """Seeded in-memory stand-ins for the (wav, hdf5) pairs train_generator reads: int16 waveforms and fp64
    WORLD-style feature matrices (feature_extract.py:343 writes fp64), lengths deliberately inconsistent so that
    both branches of _validate_length run; one utterance carries an unvoiced hole (f0 == 0)."""
    rs = np.random.RandomState(11)
    D, U = 39, synth.UPSAMPLING
    utts = []
    for u, (frames, extra) in enumerate(((23, 17), (31, -45), (18, 0), (40, 260))):
        hs, f0, _ = synth.utterance(frames, 500 + u, 1.0, D)
        raw = rs.randn(frames, D)
        raw[:, 0] = (rs.rand(frames) > 0.3)
        raw[:, 1] = f0
        if u == 2:
            raw[5, 1] = 0.0
        n = frames * U + extra
        wav = np.clip(np.round(synth.noise_waveform(n, 40 + u) * 32768), -32768, 32767).astype(np.int16)
        utts.append((wav, raw))
    mean = rs.randn(D) * 2.0
    scale = rs.rand(D) * 3.0 + 0.5
    mean[0], scale[0] = 0.0, 1.0
    return utts, mean, scale


def golden_segmenter():
    """train_generator itself (qpnet_train.py:200-335, unmodified function body) over the in-memory utterances: the
    file readers are the only stubs (wavfile.read / read_hdf5 return the arrays, check_filenames is true, the
    background-prefetch decorator is the identity).  Two passes over the list, small batch_length so that several
    segments are cut per utterance and the batch_mod1 / batch_mod2 clamps are exercised."""
    import copy, logging, types
    from numpy.matlib import repmat
    from sklearn.preprocessing import StandardScaler
    utts, mean, scale = segmenter_utterances()
    scaler = StandardScaler()
    scaler.mean_, scaler.scale_, scaler.n_features_in_ = mean, scale, mean.shape[0]
    wavfile = types.SimpleNamespace(read=lambda f: (synth.FS, utts[int(f)][0]))
    ns = {"np": np, "torch": torch, "copy": copy, "repmat": repmat, "logging": logging, "wavfile": wavfile,
          "read_hdf5": lambda f, key: utts[int(f)][1], "check_filenames": lambda l: True,
          "background": lambda max_prefetch=1: (lambda gen: gen)}
    _reference_functions("/root/reference/src/utils/utils.py", ["extend_time"], ns)
    _reference_functions("/root/reference/src/bin/qpnet_train.py",
                         ["_validate_length", "_dilated_factor", "_batch_f0", "_receptive_field", "train_generator"], ns)
    cuda = torch.cuda.is_available
    torch.cuda.is_available = lambda: False
    out = {"mean": mean, "scale": scale}
    try:
        for cfg, (bl, bs, maxlen) in enumerate(((1500, 1, 30000), (2200, 2, 3000))):
            gen = ns["train_generator"]([str(i) for i in range(len(utts))], [str(i) for i in range(len(utts))],
                                        1, 45, 15, wav_transform=lambda x: ref.encode_mu_law(x, 256),
                                        feat_transform=lambda x: scaler.transform(x), feature_type="world",
                                        dense_factor=synth.DENSE_FACTOR, batch_length=bl, batch_size=bs, max_length=maxlen,
                                        f0_threshold=0, upsampling_factor=synth.UPSAMPLING, shuffle=False)
            n = 0
            for k, (bx, bh, bt, bd, bb) in enumerate(gen):
                if k >= (9 if bs == 1 else 4):
                    break
                for name, v in (("x", bx), ("h", bh), ("t", bt), ("d", bd), ("b", bb)):
                    out[f"c{cfg}/{k}/{name}"] = v.numpy()
                n += 1
            out[f"c{cfg}/n"] = np.int64(n)
            out[f"c{cfg}/cfg"] = np.array([bl, bs, maxlen], dtype=np.int64)
    finally:
        torch.cuda.is_available = cuda
    for i, (wav, raw) in enumerate(utts):
        out[f"wav{i}"], out[f"raw{i}"] = wav, raw
    np.savez_compressed(os.path.join(HERE, "segmenter.npz"), **out)
    print("segmenter.npz:", len(out), "arrays;", [int(out[f"c{c}/n"]) for c in range(2)], "batches")


if __name__ == "__main__":
    which = sys.argv[1:] or ["indices", "forward", "generate", "mulaw", "decode", "checkpoint", "segmenter"]
    if "segmenter" in which:
        golden_segmenter()
    if "checkpoint" in which:
        golden_checkpoint()
    if "decode" in which:
        golden_decode()
    if "indices" in which:
        golden_indices()
    if "forward" in which:
        golden_forward()
    if "generate" in which:
        golden_generate()
    if "mulaw" in which:
        golden_mulaw()
