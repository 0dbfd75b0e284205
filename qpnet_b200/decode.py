"""Decode front and back end around ``QPNet.batch_fast_generate`` (reference: ``src/bin/qpnet_decode.py``).

The reference's ``decode_generator`` (qpnet_decode.py:123-209) reads one hdf5 file per utterance, scales F0, computes the
dilated factors and z-scores the features on the host in fp64, pads to a batch and uploads; its ``_decode`` loop
(311-320) µ-law-decodes every result on the host and writes a 16-bit WAV.  Here the same arithmetic runs on the device:
``qp_feat_prepare`` in front, and the generator itself writes the 16-bit PCM as its output stage
(``QpGenerateArgs.out_pcm``; ``qp_mulaw_decode_pcm16`` is the standalone form); hdf5 / file-list handling stays with the
caller, who passes the raw (T_i, D) feature matrices and the scaler statistics.
"""
from __future__ import annotations

import math
import wave

import numpy as np
import torch

from . import ops


def batch_lists(lengths, batch_size: int):
    """Utterance indices grouped like qpnet_decode.py:149-156: sort by length (stable argsort), then
    ``np.array_split`` into ``ceil(N / batch_size)`` batches."""
    idx = np.argsort(np.asarray(lengths))
    n_batch = math.ceil(len(idx) / batch_size)
    return [b.tolist() for b in np.array_split(idx, n_batch)] if n_batch else []


def prepare_batch(feats, mean, scale, fs, dense_factor=8, upsampling_factor=110, f0_factor=1.0, f0_dim_index=1,
                  extra_memory=False, device="cuda"):
    """One batch of raw feature matrices (list of (T_i, D) fp64) -> (x (B, 1) long, h (B, D, Fmax) fp32 on the device,
    n_samples_list, d) exactly as decode_generator yields them (qpnet_decode.py:158-207): d is fp64 (B, Fmax*U) when
    ``extra_memory`` is False (a device tensor here instead of a host ndarray), fp32 otherwise."""
    B = len(feats)
    D = feats[0].shape[1]
    Fmax = max(f.shape[0] for f in feats)
    raw = np.zeros((B, Fmax, D), np.float64)                      # pad_list (73-88)
    for b, f in enumerate(feats):
        raw[b, : f.shape[0]] = f
    n_frames = torch.tensor([f.shape[0] for f in feats], dtype=torch.int32)
    dev = torch.device(device)
    h, d64, d32 = ops.feat_prepare(torch.from_numpy(raw).to(dev), n_frames.to(dev),
                                   torch.as_tensor(np.asarray(mean, np.float64)).to(dev),
                                   torch.as_tensor(np.asarray(scale, np.float64)).to(dev),
                                   f0_factor, f0_dim_index, fs, dense_factor, upsampling_factor,
                                   want_f64=not extra_memory, want_f32=extra_memory)
    x = torch.full((B, 1), 128, dtype=torch.long)                 # encode_mu_law(zeros(1)) == 128 (163, 266-267)
    n_samples_list = [f.shape[0] * upsampling_factor - 1 for f in feats]   # 184
    return x, h, n_samples_list, (d32 if extra_memory else d64)


def decode(model, feats, mean, scale, fs=22050, dense_factor=8, batch_size=32, f0_factor=1.0, f0_dim_index=1,
           extra_memory=False, mode="sampling", ids=None):
    """Generate every utterance of ``feats`` (list of raw (T_i, D) matrices).  Returns ``{id: int16 PCM ndarray}``
    (ids default to the list positions), the arrays the reference writes with ``wavfile.write`` (qpnet_decode.py:315-319).

    Everything between the raw features and the PCM stays on the device: ``qp_feat_prepare`` (front end), the
    generator, whose output stage writes the 16-bit PCM next to the symbols (``QpGenerateArgs.out_pcm``), and ONE
    device-to-host copy per batch.  Every utterance draws from the Philox stream of its position in ``feats``
    (``utt_ids``), so the noise of an utterance does not depend on the batch it lands in and no two utterances of a
    corpus share a stream (the reference's global torch RNG advances across batches, qpnet.py:508-510)."""
    from . import _lib
    ids = list(range(len(feats))) if ids is None else list(ids)
    dev = next(model.parameters()).device
    qmode = {"sampling": _lib.QP_MODE_SAMPLING, "argmax": _lib.QP_MODE_ARGMAX}[mode]
    out = {}
    for group in batch_lists([f.shape[0] for f in feats], batch_size):
        x, h, n_list, d = prepare_batch([feats[i] for i in group], mean, scale, fs, dense_factor,
                                        model.upsampling_factor, f0_factor, f0_dim_index, extra_memory, dev)
        B, max_n = len(group), max(n_list)
        pcm = torch.zeros((B, max_n), dtype=torch.int16, device=dev)
        model.generate_device(x[:, -1].to(dev).contiguous(), h, d, torch.tensor(n_list, dtype=torch.int32, device=dev), max_n,
                              qmode, n_host=n_list, utt_ids=torch.tensor(group, dtype=torch.int32, device=dev), pcm_out=pcm)
        host = pcm.cpu().numpy()                                   # the only device -> host copy of the batch
        for k, n in enumerate(n_list):
            out[ids[group[k]]] = host[k, :n].copy()
    return out


def write_wav(path, fs, pcm16):
    """16-bit mono PCM, what ``scipy.io.wavfile.write(path, fs, int16 array)`` produces (qpnet_decode.py:319)."""
    with wave.open(path, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(int(fs))
        w.writeframes(np.asarray(pcm16, dtype="<i2").tobytes())
