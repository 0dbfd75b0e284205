"""CPU suite: pin the oracle (oracle/qpnet_oracle.py) to the fixtures that
tests/golden/make_golden.py produced by running the UNMODIFIED reference
(/root/reference/src/nets/qpnet.py).  Indices and symbols: exact.  Floats: 1e-5."""
import os

import numpy as np
import pytest
import torch

from oracle import qpnet_oracle as orc
from qpnet_b200 import synth
from tests import cases


def test_mulaw_known_answers():
    g = cases.load("mulaw")
    assert orc.encode_mu_law(g["x_known"]).tolist() == [0, 16, 98, 128, 157, 239, 255]
    assert np.array_equal(orc.encode_mu_law(g["x_known"]), g["enc_known"])
    assert np.array_equal(orc.encode_mu_law(g["x_rand"]), g["enc_rand"])
    np.testing.assert_allclose(orc.decode_mu_law(np.arange(256)), g["dec_all"], rtol=0, atol=1e-15)
    assert orc.decode_mu_law(128) == 0.0
    np.testing.assert_allclose(orc.decode_mu_law(0), -1.022070168, atol=1e-9)
    np.testing.assert_allclose(orc.decode_mu_law(255), 0.9784045842, atol=1e-9)


@pytest.mark.parametrize("factor", [1.0, 0.5, 1.5])
@pytest.mark.parametrize("n", [1100, 20020, 30030])
def test_indices_bit_exact(factor, n):
    g = cases.load("indices")
    frames = n // synth.UPSAMPLING
    f0 = np.stack([synth.f0_contour(frames, u) * factor for u in (0, 1)])
    d64 = np.stack([cases.d_from_f0(f) for f in f0])
    d32 = d64.astype(np.float32)
    pos = np.arange(-n, 0)
    for dil in (1, 2, 4, 8):
        key = f"f{factor}_n{n}_d{dil}"
        assert np.array_equal(orc.tf_index_f32(d32, dil) - pos, g[key + "_tf32"])
        assert np.array_equal(orc.tf_index_f64(d64, dil) - pos, g[key + "_tf64"])
        assert np.array_equal(orc.gen_index_f32(d32, dil), g[key + "_g32"])
        assert np.array_equal(orc.gen_index_f64(d64, dil), g[key + "_g64"])


def test_tf_index_depends_on_segment_geometry():
    """SURVEY hard part 2: the fp32 rounding of the SUM makes idx depend on n."""
    g = cases.load("indices")
    naive_diff = 0
    for dil in (1, 2, 4, 8):
        a = g[f"f1.0_n30030_d{dil}_tf32"].astype(np.int64)     # look-back = idx - pos
        b = g[f"f1.0_n30030_d{dil}_g32"].astype(np.int64)
        naive_diff += int((a != b).sum())
    assert naive_diff > 0


@pytest.mark.parametrize("name", list(cases.FORWARD_CASES))
def test_forward_logits(name):
    g = cases.load("forward")
    kw, a, p, x, h, d, t, bl = cases.forward_inputs(name)
    assert np.array_equal(x.numpy()[0], g[f"{name}/x"])
    with torch.no_grad():
        logits = orc.forward(a, p, x, h, d, bl)[0].numpy()
    np.testing.assert_allclose(logits, g[f"{name}/logits"], rtol=0, atol=2e-5)


@pytest.mark.parametrize("name", [n for n, c in cases.FORWARD_CASES.items() if c[-1]])
def test_forward_grads(name):
    g = cases.load("forward")
    kw, a, p, x, h, d, t, bl = cases.forward_inputs(name)
    assert np.array_equal(t.numpy()[0], g[f"{name}/target"])
    p = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    logits = orc.forward(a, p, x, h, d, bl)
    loss = torch.nn.functional.cross_entropy(logits.reshape(-1, a.Q), t.reshape(-1))
    loss.backward()
    np.testing.assert_allclose(loss.item(), float(g[f"{name}/loss"]), atol=1e-5)
    last = f"resA_1x1.{len(a.dilA) - 1}"
    for k, v in p.items():
        ref = g[f"{name}/grad/{k}"]
        if k.startswith(last):                      # dead projection: reference grad is None (C7)
            assert ref.size == 0
            assert v.grad is None or float(v.grad.abs().max()) == 0.0
            continue
        scale = max(float(np.abs(ref).max()), 1e-6)
        np.testing.assert_allclose(v.grad.numpy(), ref, rtol=0, atol=2e-4 * scale + 1e-7, err_msg=k)


@pytest.mark.parametrize("name", list(cases.GENERATE_CASES))
def test_generate_symbols(name):
    g = cases.load("generate")
    kw, a, p, x, h, d, n_list, mode, xm = cases.generate_inputs(name)
    uni = torch.from_numpy(g[f"{name}/uniforms"])
    with torch.no_grad():
        res = orc.generate(a, p, x, h, n_list, d, mode=mode, uniforms=uni, f64_index=not xm)
    for b, r in enumerate(res):
        ref = g[f"{name}/sym{b}"].astype(np.int64)
        assert len(r) == n_list[b] == len(ref)
        match = float((r == ref).mean())
        # identical algorithm, different fp32 summation order: allow a late divergence only
        first = int(np.argmax(r != ref)) if match < 1.0 else len(ref)
        assert match == 1.0 or first > 50, (name, b, match, first)
        assert match > 0.5 or first > 50


def test_generate_teacher_forced_logits_match_forward():
    """State-machine check independent of chaos: feed fixed symbols to the generator and
    compare its per-step logits with the teacher-forced forward on the same symbols."""
    kw, a, p, x, h, d, n_list, mode, xm = cases.generate_inputs("small_sampling")
    B = len(n_list)
    steps = 300
    rs = np.random.RandomState(3)
    forced = torch.from_numpy(rs.randint(0, a.Q, size=(B, steps)))
    lg = []
    with torch.no_grad():
        orc.generate(a, p, x, h, n_list, d, mode="argmax", force=forced, logits_out=lg,
                     max_steps=steps)
    lg = torch.stack(lg, dim=1)                               # (B, steps, Q)
    M = int(np.ceil(d).max())
    R = a.rfA * M + a.rfF + a.rfC
    for b in range(B):
        # teacher-forced input: [half]*R, seed, forced[0..steps-2]; logits for positions 0..steps-1
        xs = torch.cat([torch.full((R + 1,), a.Q // 2), forced[b, : steps - 1]])
        T = xs.numel()
        pad = (-T) % a.U
        xs = torch.cat([torch.full((pad,), a.Q // 2), xs])
        hup_needed = xs.numel()
        # aux/d for the pad region: replicate frame 0 / d=1 is only exact if frame aligned;
        # build sample-rate tensors directly through a unit upsampler equivalent
        hfull = orc._upsample(p, h[b])                        # (T_up, A)
        hrows = torch.cat([hfull[:1].expand(R + pad, -1), hfull[:steps]])
        drow = torch.cat([torch.ones(R + pad, dtype=torch.float64), torch.from_numpy(d[b, :steps])])
        with torch.no_grad():
            lt = _forward_rows(a, p, xs, hrows, drow.float(), steps, M)
        np.testing.assert_allclose(lg[b].numpy(), lt.numpy(), atol=2e-4)


def _forward_rows(a, p, x, hup, d, bl, M):
    """forward_one with sample-rate aux rows given directly (test helper).  Uses the
    GENERATION index flavour (no position term) so that it mirrors the generator."""
    rfA = a.rfA * M
    R = rfA + a.rfF + a.rfC
    cur = orc._embed(p, x[-R - bl:])
    hup = hup[-(R - 1 + bl):]
    d = d[-(R - 1 + bl):]
    skips = 0
    for i, dil in enumerate(a.dilF):
        L = cur.shape[0]
        xp, xc = cur[: L - dil], cur[dil:]
        _, skip, res = orc._gate(p, "F", i, xp, xc, hup[-(L - dil):])
        cur = res + xc
        skips = skips + skip[-bl:]
    for i, dil in enumerate(a.dilA):
        shift = dil * M
        L = cur.shape[0]
        n = L - shift
        k = -torch.from_numpy(orc.gen_index_f64(d[-n:].double().numpy(), dil)).long()
        pos = torch.arange(shift, L) - k
        xc, xp = cur[shift:], cur[pos]
        _, skip, res = orc._gate(p, "A", i, xp, xc, hup[-n:])
        cur = res + xc
        skips = skips + skip[-bl:]
    return orc._head(p, skips)


# ------------------------------------------------------------------ decode front / back end
def test_decode_frontend_and_pcm_match_reference_functions():
    """oracle.decode_frontend / decode_pcm16 against fixtures produced by the reference's own pad_list,
    _dilated_factor, _batch_f0, extend_time, decode_mu_law and sklearn's StandardScaler (make_golden.py: golden_decode)."""
    g = cases.load("decode")
    for fac in (1.0, 0.5, 1.5):
        feats = [g[f"f{fac}/raw{u}"] for u in range(3)]
        h, d, ns = orc.decode_frontend(feats, g["mean"], g["scale"], synth.FS, synth.DENSE_FACTOR, synth.UPSAMPLING, fac, 1)
        assert h.dtype == np.float32 and np.array_equal(h, g[f"f{fac}/h"])
        assert np.array_equal(d, g[f"f{fac}/d"])
        assert ns == [f.shape[0] * synth.UPSAMPLING - 1 for f in feats]
    assert np.array_equal(orc.decode_pcm16(np.arange(256)), g["pcm_all_symbols"])


def test_decode_batch_lists_follow_the_reference_split():
    """qpnet_decode.py:149-156: argsort by length, then np.array_split into ceil(N / batch_size) batches."""
    from qpnet_b200.decode import batch_lists          # host logic only (no CUDA call)
    lengths = [50, 10, 30, 20, 40, 60, 5]
    got = batch_lists(lengths, 3)
    assert got == [[6, 1, 3], [2, 4], [0, 5]]
    assert batch_lists([], 4) == []


def test_train_segmenter_oracle_vs_reference_train_generator():
    """oracle.train_segments against batches cut by the reference's own train_generator (qpnet_train.py:200-335, run
    over in-memory utterances by tests/golden/make_golden.py::golden_segmenter): every array bit exact -- symbols,
    z-scored features, dilated factors and the per-batch segment lengths, for both configurations."""
    g = np.load(os.path.join(cases.GOLDEN, "segmenter.npz"))
    utts = [(g[f"wav{i}"], g[f"raw{i}"]) for i in range(4)]
    for c in range(2):
        bl, bs, ml = [int(v) for v in g[f"c{c}/cfg"]]
        gen = orc.train_segments(utts, g["mean"], g["scale"], 1, 45, 15, synth.FS, synth.DENSE_FACTOR, bl, bs, ml, 0,
                                 synth.UPSAMPLING, 256, passes=3)
        n = int(g[f"c{c}/n"])
        assert n >= 4
        lengths = set()
        for k in range(n):
            got = next(gen)
            for name, v in zip("xhtdb", got):
                ref = g[f"c{c}/{k}/{name}"]
                assert v.shape == ref.shape and v.dtype == ref.dtype, (c, k, name, v.dtype, ref.dtype)
                np.testing.assert_array_equal(v, ref)
            lengths.add(int(got[4][0]))
        assert len(lengths) >= 2          # the receptive field (hence the segment length) really changed along the stream

