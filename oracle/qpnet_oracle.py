"""CPU ORACLE for the QPNet hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this file.  Nothing under ``qpnet_b200/``
imports it, and the product path raises when its CUDA library is missing.

It restates, on the CPU, what ``/root/reference/src/nets/qpnet.py`` computes on
the path this repository replaces.  Integer index arithmetic is restated in numpy
with explicit float32 / float64 rounding (bit-exact by construction); the floating
point stack is restated with plain torch CPU fp32 tensor algebra (matmul on
time-major activations, no Conv1d modules), which keeps autograd available so the
same restatement is the oracle for gradients.

PARITY PIN: the reference ships no tests and no golden vectors, so
this oracle is pinned against outputs of the *reference itself* executed in the
authoring container: ``tests/golden/make_golden.py`` imports the unmodified
``/root/reference/src/nets/qpnet.py`` and writes the fixtures under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks every function below
against them (indices exact, floats to 1e-5).  Third-party arithmetic underneath the
reference (ATen / oneDNN convolutions, torch 2.11 here, "PyTorch 1.3" pinned by
``README.md:20``) is not vendored and not pinned by any reference-owned vector.

Every function cites the reference lines it follows.
"""
from __future__ import annotations

import numpy as np
import torch

# --------------------------------------------------------------------------------------
# mu-law (qpnet.py:22-45)
# --------------------------------------------------------------------------------------

def encode_mu_law(x, mu=256):
    """qpnet.py:22-32 -- float waveform in [-1,1] -> int64 symbols in [0, mu)."""
    m = mu - 1
    x = np.asarray(x, dtype=np.float64)
    compressed = np.sign(x) * np.log1p(m * np.abs(x)) / np.log1p(m)
    return np.floor((compressed + 1.0) / 2.0 * m + 0.5).astype(np.int64)


def decode_mu_law(y, mu=256):
    """qpnet.py:34-45 -- note the asymmetric -0.5 offset of the reference."""
    m = mu - 1
    y = np.asarray(y, dtype=np.float64)
    c = (y - 0.5) / m * 2.0 - 1.0
    return np.sign(c) / m * (np.power(1.0 + m, np.abs(c)) - 1.0)


# --------------------------------------------------------------------------------------
# caller-side F0 -> dilated factor (qpnet_train.py:147-165, qpnet_decode.py:90-108,
# utils/utils.py:216-235)
# --------------------------------------------------------------------------------------

def dilated_factor(f0, fs, dense_factor, f0_threshold=None):
    """d = ((1.0*fs)/f0)/dense in float64; f0 == 0 is replaced by fs/dense (-> d = 1).

    ``f0_threshold`` (train side only, qpnet_train.py:178) clamps F0 from below first.
    """
    f = np.array(f0, dtype=np.float64, copy=True)
    if f0_threshold is not None:
        f[f < f0_threshold] = f0_threshold
    f[f == 0] = fs / dense_factor
    d = np.ones(f.shape) * fs
    d /= f
    d /= dense_factor
    assert np.all(d > 0)
    return d


def extend_time(v, upsampling_factor):
    """utils.py:216-235 -- repeat every frame value ``upsampling_factor`` times."""
    return np.repeat(np.asarray(v), upsampling_factor, axis=0)


# --------------------------------------------------------------------------------------
# dilation-index builders (qpnet.py:592-624).  Three normative flavours.
# --------------------------------------------------------------------------------------

def tf_index_f32(d, dilation):
    """qpnet.py:594-600 (tensor path): rint_f32( (-d*dil)_f32 + float(t-n) ), int64.

    ``d``: (B, n) float32.  Result is negative, relative to the END of the layer input.
    """
    d = np.asarray(d, dtype=np.float32)
    n = d.shape[-1]
    pos = np.arange(-n, 0).astype(np.float32)
    prod = (-d) * np.float32(dilation)              # one fp32 rounding
    s = (prod + pos).astype(np.float32)             # second fp32 rounding
    return np.rint(s).astype(np.int64)              # half-to-even, like torch.round


def tf_index_f64(d, dilation):
    """qpnet.py:606-609 (numpy path, generation priming): same in float64 -> int32."""
    d = np.asarray(d, dtype=np.float64)
    n = d.shape[-1]
    s = (-d * dilation) + np.arange(-n, 0)
    return np.int32(np.round(s))


def gen_index_f32(d, dilation):
    """qpnet.py:615-617 (extra_memory=True): rint_f32(-d*dil), int64. No position term."""
    d = np.asarray(d, dtype=np.float32)
    return np.rint((-d) * np.float32(dilation)).astype(np.int64)


def gen_index_f64(d, dilation):
    """qpnet.py:621-622 (extra_memory=False, the shipped default): float64 -> int32."""
    d = np.asarray(d, dtype=np.float64)
    return np.int32(np.round(-d * dilation))


# --------------------------------------------------------------------------------------
# model description + parameters (qpnet.py:174-237, initialize 47-58)
# --------------------------------------------------------------------------------------

class Arch:
    """Hyper-parameters of qpnet.py:174-199 and the derived receptive fields."""

    def __init__(self, n_quantize=256, n_aux=39, n_resch=512, n_skipch=256,
                 dilationF_depth=4, dilationF_repeat=3, dilationA_depth=4,
                 dilationA_repeat=1, kernel_size=2, upsampling_factor=110):
        assert kernel_size == 2
        self.Q, self.A, self.C, self.S = n_quantize, n_aux, n_resch, n_skipch
        self.U = upsampling_factor
        self.dilF = [2 ** i for i in range(dilationF_depth)] * dilationF_repeat
        self.dilA = [2 ** i for i in range(dilationA_depth)] * dilationA_repeat
        self.rfC = 1
        self.rfF = sum(self.dilF)
        self.rfA = sum(self.dilA)


def state_dict_spec(a: Arch):
    """Ordered (name, shape) list == ``QPNet(...).state_dict()`` of the reference."""
    C, S, Q, A, U = a.C, a.S, a.Q, a.A, a.U
    spec = [("causal.conv.weight", (C, Q, 2)), ("causal.conv.bias", (C,)),
            ("upsampling.conv.weight", (1, 1, 1, U)), ("upsampling.conv.bias", (1,))]
    nF, nA = len(a.dilF), len(a.dilA)
    for grp in ("dilF_sigmoid", "dilF_tanh"):
        for i in range(nF):
            spec += [(f"{grp}.{i}.conv.weight", (C, C, 2)), (f"{grp}.{i}.conv.bias", (C,))]
    for grp in ("auxF_1x1_sigmoid", "auxF_1x1_tanh"):
        for i in range(nF):
            spec += [(f"{grp}.{i}.weight", (C, A, 1)), (f"{grp}.{i}.bias", (C,))]
    for i in range(nF):
        spec += [(f"skipF_1x1.{i}.weight", (S, C, 1)), (f"skipF_1x1.{i}.bias", (S,))]
    for i in range(nF):
        spec += [(f"resF_1x1.{i}.weight", (C, C, 1)), (f"resF_1x1.{i}.bias", (C,))]
    for grp in ("dilA_sigmoid", "dilA_tanh"):
        for i in range(nA):
            spec += [(f"{grp}.{i}.convC.weight", (C, C, 1)), (f"{grp}.{i}.convC.bias", (C,)),
                     (f"{grp}.{i}.convP.weight", (C, C, 1)), (f"{grp}.{i}.convP.bias", (C,))]
    for grp in ("auxA_1x1_sigmoid", "auxA_1x1_tanh"):
        for i in range(nA):
            spec += [(f"{grp}.{i}.weight", (C, A, 1)), (f"{grp}.{i}.bias", (C,))]
    for i in range(nA):
        spec += [(f"skipA_1x1.{i}.weight", (S, C, 1)), (f"skipA_1x1.{i}.bias", (S,))]
    for i in range(nA):
        spec += [(f"resA_1x1.{i}.weight", (C, C, 1)), (f"resA_1x1.{i}.bias", (C,))]
    spec += [("conv_post_1.weight", (S, S, 1)), ("conv_post_1.bias", (S,)),
             ("conv_post_2.weight", (Q, S, 1)), ("conv_post_2.bias", (Q,))]
    return spec


def init_params(a: Arch, seed: int, bias_std: float = 0.0):
    """Random parameters in the spirit of ``initialize`` (qpnet.py:47-58): Xavier-uniform
    conv weights, unit upsampler.  ``bias_std`` > 0 draws N(0, bias_std) biases (and a
    perturbed upsampler) so that bias paths are exercised -- used for parity sets only.
    NOT bit-identical to torch's initialiser; parity tests load *these* tensors into
    both sides instead.
    """
    g = torch.Generator().manual_seed(seed)
    p = {}
    for name, shape in state_dict_spec(a):
        if name.startswith("upsampling"):
            base = 1.0 if name.endswith("weight") else 0.0
            t = torch.full(shape, base)
            if bias_std > 0:
                t = t + bias_std * torch.randn(shape, generator=g)
        elif name.endswith("bias"):
            t = bias_std * torch.randn(shape, generator=g) if bias_std > 0 else torch.zeros(shape)
        else:
            fan_out, fan_in = shape[0] * shape[2], shape[1] * shape[2]
            lim = (6.0 / (fan_in + fan_out)) ** 0.5
            t = (torch.rand(shape, generator=g) * 2 - 1) * lim
        p[name] = t.float()
    return p


# --------------------------------------------------------------------------------------
# building blocks on time-major activations  (rows = time, cols = channels)
# --------------------------------------------------------------------------------------

def _embed(p, x):
    """OneHot + CausalConv1d k=2, no padding (qpnet.py:60-79,110-132,561-564):
    y[t] = W[:, x[t], 0] + W[:, x[t+1], 1] + b.   x: (T,) int64 -> (T-1, C)."""
    w = p["causal.conv.weight"]
    x = x % w.shape[1]
    return w[:, x[:-1], 0].t() + w[:, x[1:], 1].t() + p["causal.conv.bias"]


def _upsample(p, h):
    """UpSampling (qpnet.py:134-158): one shared U-tap transposed conv, stride U.
    h: (A, F) -> (F*U, A) time-major."""
    w = p["upsampling.conv.weight"].reshape(-1)
    b = p["upsampling.conv.bias"].reshape(())
    up = h.t().unsqueeze(1) * w.view(1, -1, 1) + b       # (F, U, A)
    return up.reshape(-1, h.shape[0])


def _gate(p, kind, i, x_past, x_cur, h_rows):
    """Gated unit of one residual block -> (z, skip, res_out_without_residual).

    fixed  (qpnet.py:657-670): taps 0/1 of the k=2 dilated conv are past/current.
    adaptive (qpnet.py:626-640, 89-108): convP / convC 1x1 convs, each with a bias.
    """
    pre = []
    for g in ("sigmoid", "tanh"):
        if kind == "F":
            w = p[f"dilF_{g}.{i}.conv.weight"]
            s = x_past @ w[:, :, 0].t() + x_cur @ w[:, :, 1].t() + p[f"dilF_{g}.{i}.conv.bias"]
        else:
            s = (x_cur @ p[f"dilA_{g}.{i}.convC.weight"][:, :, 0].t() + p[f"dilA_{g}.{i}.convC.bias"]
                 + x_past @ p[f"dilA_{g}.{i}.convP.weight"][:, :, 0].t() + p[f"dilA_{g}.{i}.convP.bias"])
        s = s + h_rows @ p[f"aux{kind}_1x1_{g}.{i}.weight"][:, :, 0].t() + p[f"aux{kind}_1x1_{g}.{i}.bias"]
        pre.append(s)
    z = torch.sigmoid(pre[0]) * torch.tanh(pre[1])
    skip = z @ p[f"skip{kind}_1x1.{i}.weight"][:, :, 0].t() + p[f"skip{kind}_1x1.{i}.bias"]
    res = z @ p[f"res{kind}_1x1.{i}.weight"][:, :, 0].t() + p[f"res{kind}_1x1.{i}.bias"]
    return z, skip, res


def _head(p, skip_sum):
    """_postprocess (qpnet.py:566-571): relu -> 1x1 -> relu -> 1x1."""
    y = torch.relu(skip_sum) @ p["conv_post_1.weight"][:, :, 0].t() + p["conv_post_1.bias"]
    return torch.relu(y) @ p["conv_post_2.weight"][:, :, 0].t() + p["conv_post_2.bias"]


# --------------------------------------------------------------------------------------
# teacher-forced forward (qpnet.py:239-312) -- one batch element at a time
# --------------------------------------------------------------------------------------

def forward_one(a: Arch, p, x, h, d, bl, taps=None):
    """Logits (bl, Q) for ONE utterance/segment.

    x: (T,) int64, h: (A, T/U) fp32, d: (T,) fp32, all aligned from the END like the
    reference slices them (x[:, -R-bl:], h[:, :, hindex:], d[:, hindex:]).
    The reference gathers every batch element's past taps from element 0
    (qpnet.py:250, caveat C1), so parity is defined against B=1 calls.
    ``taps``: optional dict that receives per-layer activations.
    """
    x = torch.as_tensor(x, dtype=torch.int64)
    h = torch.as_tensor(h, dtype=torch.float32)
    d = torch.as_tensor(d, dtype=torch.float32)
    M = int(torch.max(torch.ceil(d)))                       # qpnet.py:255 (whole d passed)
    rfA = a.rfA * M
    R = rfA + a.rfF + a.rfC                                  # qpnet.py:261
    cur = _embed(p, x[-R - bl:])                             # (rfA+rfF+bl, C)
    hup = _upsample(p, h)                                    # (T, A)
    skips = 0
    hindex = -(rfA + a.rfF + bl)                             # qpnet.py:269
    for i, dil in enumerate(a.dilF):
        hindex += dil
        L = cur.shape[0]
        x_past, x_cur = cur[: L - dil], cur[dil:]
        z, skip, res = _gate(p, "F", i, x_past, x_cur, hup[hindex:] if hindex < 0 else hup[:0])
        cur = res + x_cur
        skips = skips + skip[-bl:]
        if taps is not None:
            taps[f"F{i}"] = cur
    hindex = -(rfA + bl)                                     # qpnet.py:287
    for i, dil in enumerate(a.dilA):
        shift = dil * M
        hindex += shift
        L = cur.shape[0]
        dd = d[hindex:] if hindex < 0 else d[:0]
        idx = torch.from_numpy(tf_index_f32(dd.numpy(), dil))        # qpnet.py:292-293
        assert int(abs(idx.min())) <= L                      # qpnet.py:294
        x_cur = cur[shift:]
        x_past = cur[L + idx]                                # negative index from the end
        z, skip, res = _gate(p, "A", i, x_past, x_cur, hup[hindex:] if hindex < 0 else hup[:0])
        cur = res + x_cur
        skips = skips + skip[-bl:]
        if taps is not None:
            taps[f"A{i}"] = cur
            taps[f"idxA{i}"] = idx
    return _head(p, skips)


def forward(a: Arch, p, x, h, d, bl):
    """Batched wrapper: per-element ``forward_one`` stacked to (B, bl, Q)."""
    return torch.stack([forward_one(a, p, x[b], h[b], d[b], bl) for b in range(len(x))])


# --------------------------------------------------------------------------------------
# autoregressive generation (qpnet.py:314-559)
# --------------------------------------------------------------------------------------

def inverse_cdf(logits, u):
    """Shared deterministic sampler: softmax -> cumulative sum -> first k with
    cdf[k] > u (clamped).  Used on BOTH sides of the free-running comparison in place
    of ``Categorical.sample`` (qpnet.py:508-510), whose RNG stream cannot be shared."""
    pr = torch.softmax(logits, dim=-1)
    cdf = torch.cumsum(pr, dim=-1)
    k = torch.searchsorted(cdf, u.reshape(-1, 1).to(cdf.dtype).contiguous(), right=True).reshape(-1)
    return torch.clamp(k, max=logits.shape[-1] - 1)


def generate(a: Arch, p, seed_x, h, n_samples_list, d, mode="sampling", uniforms=None,
             f64_index=True, force=None, logits_out=None, max_steps=None):
    """Restatement of ``batch_fast_generate``; returns list of int64 arrays in INPUT
    order (the reference returns ascending-length order and mutates the caller's list,
    qpnet.py:527-559 / caveat C3 -- that host behaviour is mirrored by the product
    wrapper, not here).

    seed_x: (B, 1) int64; h: (B, A, Fmax) fp32 zero-padded; d: (B, U*Fmax) float64
    (extra_memory=False flavour, ``f64_index=True``) or fp32.
    Priming (qpnet.py:355-440): x is left-padded with Q/2, h replicate-padded, d with
    1.0, so every layer output is constant over the pad region; each FIFO therefore
    starts filled with that constant, which is computed here on a length-1 time axis.
    Step i consumes h_up[:, i], d[i] and the previous symbol (qpnet.py:446-516).
    FIFO semantics: fixed layers look back exactly ``dil`` steps; adaptive layers look
    back k = -round(-d[i]*dil) steps into the previous layer's FIFO of depth dil*M,
    k == 0 selecting the OLDEST entry (python index 0, caveat C4).
    ``force``: optional (B, steps) symbols fed back instead of the drawn ones
    (teacher-forced generator check).  ``uniforms``: (B, steps) for mode="sampling".
    Runs on the device the parameters live on (CPU for parity; bench.py also times this port with
    CUDA tensors as the "reference's own eager-PyTorch GPU path on this box" row).
    """
    B = len(n_samples_list)
    dev = p["causal.conv.weight"].device
    h = torch.as_tensor(h, dtype=torch.float32).to(dev)
    if uniforms is not None:
        uniforms = torch.as_tensor(uniforms).to(dev)
    if force is not None:
        force = torch.as_tensor(force).to(dev)
    d_np = np.asarray(d, dtype=np.float64 if f64_index else np.float32)
    T = max(n_samples_list) if max_steps is None else max_steps
    M = int(np.nanmax(np.ceil(d_np)))                        # qpnet.py:347-350
    half = a.Q // 2
    hup = torch.stack([_upsample(p, h[b]) for b in range(B)])          # (B, T_up, A)
    gidx = [(gen_index_f64 if f64_index else gen_index_f32)(d_np, dil) for dil in a.dilA]
    nF, nA = len(a.dilF), len(a.dilA)
    dils = list(a.dilF) + list(a.dilA)
    depth = list(a.dilF) + [dl * M for dl in a.dilA]         # look-back bound per layer input
    # ---- priming: constant signal ------------------------------------------------------
    sym = torch.full((B, 2), half, dtype=torch.int64, device=dev)
    x0 = torch.stack([_embed(p, sym[b])[0] for b in range(B)])        # (B, C)
    hist = []                                                # hist[l]: list of (B, C), newest last
    cur = x0
    h0 = hup[:, 0]
    for l in range(nF + nA):
        hist.append([cur] * depth[l])
        kind, i = ("F", l) if l < nF else ("A", l - nF)
        _, _, res = _gate(p, kind, i, cur, cur, h0)
        cur = res + cur
    # ---- sample loop -------------------------------------------------------------------
    prev = torch.full((B,), half, dtype=torch.int64, device=dev)
    new = torch.as_tensor(seed_x, dtype=torch.int64).reshape(B, -1)[:, -1].clone().to(dev)
    out = torch.zeros((B, T), dtype=torch.int64, device=dev)
    for i in range(T):
        pair = torch.stack([prev, new], dim=1)
        cur = torch.stack([_embed(p, pair[b])[0] for b in range(B)])
        hC = hup[:, min(i, hup.shape[1] - 1)]
        skips = 0
        for l in range(nF + nA):
            if l < nF:
                past = hist[l][-dils[l]]
                kind, j = "F", l
            else:
                j = l - nF
                k = -gidx[j][:, min(i, gidx[j].shape[1] - 1)]
                rows = []
                for b in range(B):
                    kb = int(k[b])
                    pos = depth[l] - kb if kb > 0 else 0     # depth+idx ; idx==0 -> oldest
                    rows.append(hist[l][pos][b])
                past = torch.stack(rows)
                kind = "A"
            _, skip, res = _gate(p, kind, j, past, cur, hC)
            hist[l].append(cur)
            hist[l] = hist[l][-depth[l]:]
            cur = res + cur
            skips = skips + skip
        logits = _head(p, skips)
        if logits_out is not None:
            logits_out.append(logits.clone())
        if mode == "sampling":
            s = inverse_cdf(logits, uniforms[:, i])
        elif mode == "argmax":
            s = logits.argmax(-1)
        else:
            raise ValueError("mode should be sampling or argmax")
        out[:, i] = s
        prev = new
        new = s if force is None else torch.as_tensor(force[:, i], dtype=torch.int64)
    out = out.cpu()
    return [out[b, : min(n_samples_list[b], T)].numpy() for b in range(B)]


# ------------------------------------------------------------------ decode front / back end (next row (f)1)
def decode_pad_list(batch_list, pad_value=0.0):
    """qpnet_decode.py:73-88."""
    maxlen = max(b.shape[0] for b in batch_list)
    out = np.zeros((len(batch_list), maxlen, batch_list[0].shape[-1]))
    for i, b in enumerate(batch_list):
        out[i, : b.shape[0]] = b
    return out


def decode_frontend(feats, mean, scale, fs, dense_factor, upsampling_factor, f0_factor, f0_dim_index):
    """What decode_generator yields for one batch of raw (T_i, D) matrices (qpnet_decode.py:158-200, 268-269):
    h (B, D, Fmax) float32, d (B, Fmax*U) float64, n_samples_list."""
    hs, ds, ns = [], [], []
    for f in feats:
        h = np.array(f, dtype=np.float64, copy=True)
        h[:, f0_dim_index] = h[:, f0_dim_index] * f0_factor                       # 172-173
        d = dilated_factor(h[:, f0_dim_index].copy(order="C"), fs, dense_factor)  # 174, 90-108
        d = extend_time(d, upsampling_factor)                                     # 175
        h = (h - np.asarray(mean, np.float64)) / np.asarray(scale, np.float64)    # StandardScaler.transform (268-269)
        hs.append(h)
        ds.append(d[:, None])
        ns.append(f.shape[0] * upsampling_factor - 1)                             # 184
    h_pad = np.transpose(decode_pad_list(hs).astype(np.float32), (0, 2, 1))       # 187, 190
    d_pad = decode_pad_list(ds)[:, :, 0]                                          # 188, 194
    return h_pad, d_pad, ns


def decode_pcm16(samples, mu=256):
    """qpnet_decode.py:315-318: decode_mu_law -> * 32768 -> clip -> int16."""
    wav = decode_mu_law(np.asarray(samples), mu)
    return np.clip(wav * 32768, -32768, 32767).astype(np.int16)


# --------------------------------------------------------------------------------------
# training segmenter (qpnet_train.py:119-145 _validate_length, 181-199 _receptive_field, 242-335 train_generator)
# --------------------------------------------------------------------------------------

def validate_length(x, h, upsampling):
    """qpnet_train.py:119-145 with an upsampling factor: trim the waveform to whole frames, or drop the frames the
    waveform does not reach (one more than strictly needed, like the reference)."""
    if len(x) > len(h) * upsampling:
        x = x[: len(h) * upsampling]
    if len(x) < len(h) * upsampling:
        short = len(h) * upsampling - len(x)
        h = h[: len(h) - (short // upsampling + 1)]
        x = x[: len(h) * upsampling]
    return x, h


def train_segments(utterances, mean, scale, rf_causal, rf_fixed, rf_adaptive, fs, dense_factor, batch_length, batch_size,
                   max_length, f0_threshold, upsampling, n_quantize=256, passes=1):
    """CPU restatement of the streaming segmenter: ``utterances`` = [(int16 waveform, fp64 (frames, D) features)] in
    file order (no shuffle).  Yields (x, h, t, d, b) numpy batches exactly like qpnet_train.py:242-335: buffers grow by
    one utterance, the receptive field follows the largest dilated factor still buffered (181-199), the segment
    length is clamped by max_length and rounded down to whole frames (268-284), segments advance by the batch length
    (310-316)."""
    xb = np.empty(0, np.float32)
    hb = np.empty((0, len(mean)), np.float64)
    db = np.empty(0, np.float64)
    bx, bh, bt, bd, bb = [], [], [], [], []
    for _ in range(passes):
        for wav, raw in utterances:
            x = np.asarray(wav, np.float32) / 32768
            x, h = validate_length(x, np.asarray(raw, np.float64), upsampling)
            f0 = h[:, 1].copy()
            f0[f0 < f0_threshold] = f0_threshold
            d = extend_time(dilated_factor(f0, fs, dense_factor), upsampling)
            xb, hb, db = np.concatenate([xb, x]), np.concatenate([hb, h]), np.concatenate([db, d])
            rf = int(rf_fixed + rf_adaptive * int(np.ceil(np.nanmax(db))) + rf_causal)
            bl = batch_length - max(rf + batch_length - max_length, 0)
            bl -= (rf + bl) % upsampling
            h_bs = (rf + bl) // upsampling
            x_bs = h_bs * upsampling + 1
            want = batch_size - len(bx)
            while len(hb) > want * h_bs and len(xb) > want * x_bs:
                sym = encode_mu_law(xb[:x_bs], n_quantize)
                hz = ((hb[:h_bs] - mean) / scale).astype(np.float32)
                bx.append(sym[:-1]); bt.append(sym[1:]); bh.append(hz.T); bd.append(db[: x_bs - 1].astype(np.float32))
                bb.append(bl)
                want -= 1
                xb, hb, db = xb[bl // upsampling * upsampling:], hb[bl // upsampling:], db[bl // upsampling * upsampling:]
                if len(bx) == batch_size:
                    yield np.stack(bx), np.stack(bh), np.stack(bt), np.stack(bd), np.array(bb)
                    bx, bh, bt, bd, bb = [], [], [], [], []
                    want = batch_size

