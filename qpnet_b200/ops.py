"""Thin functional wrappers over the C ABI: torch tensors in, torch tensors out.

torch is used for device memory and streams only; every computation below is a kernel
of libqpnet_b200.so.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import check, lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts):
    for t in ts:
        if not t.is_cuda:
            raise RuntimeError("qpnet_b200 runs on a CUDA (sm_100a) device only; got a CPU tensor")


def max_ceil(d: torch.Tensor) -> int:
    """int(max(ceil(d))) -- qpnet.py:255 / 347-350 (one device->host sync, like the reference)."""
    _need_cuda(d)
    d = d.contiguous()
    out = torch.empty(1, dtype=torch.int32, device=d.device)
    fn = lib.qp_max_ceil_f64 if d.dtype == torch.float64 else lib.qp_max_ceil_f32
    if d.dtype not in (torch.float32, torch.float64):
        raise TypeError("dilated factors must be float32 or float64")
    check(fn(d.data_ptr(), d.numel(), out.data_ptr(), _stream()))
    return int(out.item())


def _ld(t: torch.Tensor) -> int:
    """Leading stride of a contiguous 2-D tensor.  (For a single row torch keeps whatever stride the tensor was
    created with -- possibly 0 -- so do not trust ``stride(0)`` there.)"""
    return t.shape[1] if t.shape[0] <= 1 else t.stride(0)


def dilated_index(d: torch.Tensor, dilation: int, flavour: str) -> torch.Tensor:
    """Bit-exact index builders of qpnet.py:592-624.

    flavour: 'tf_f32' | 'tf_f64' | 'gen_f32' | 'gen_f64' (SURVEY.md caveat C2)."""
    _need_cuda(d)
    assert d.dim() == 2
    B, n = d.shape
    if flavour in ("tf_f32", "gen_f32"):
        d = d.to(torch.float32).contiguous()
        out = torch.empty((B, n), dtype=torch.int64, device=d.device)
    elif flavour in ("tf_f64", "gen_f64"):
        d = d.to(torch.float64).contiguous()
        out = torch.empty((B, n), dtype=torch.int32, device=d.device)
    else:
        raise ValueError(flavour)
    fn = getattr(lib, "qp_index_" + flavour)
    empty = d.numel() == 0                      # empty / ragged edge case: still goes through the C ABI
    check(fn(None if empty else d.data_ptr(), B, n, n if empty else _ld(d), dilation,
             None if empty else out.data_ptr(), _stream()))
    return out


def f0_to_dilated(f0: torch.Tensor, fs: float, dense_factor: float, upsampling: int, f0_floor: float = -1.0,
                  want_f64: bool = True, want_f32: bool = True):
    """Frame-level F0 (B, F) fp64 -> per-sample dilated factors (B, F*U) (qpnet_train.py:147-179,
    qpnet_decode.py:90-120, utils.py:216-235)."""
    _need_cuda(f0)
    f0 = f0.to(torch.float64).contiguous()
    B, F = f0.shape
    d64 = torch.empty((B, F * upsampling), dtype=torch.float64, device=f0.device) if want_f64 else None
    d32 = torch.empty((B, F * upsampling), dtype=torch.float32, device=f0.device) if want_f32 else None
    check(lib.qp_f0_to_dilated(f0.data_ptr(), B, F, float(fs), float(dense_factor), upsampling, float(f0_floor),
                               d64.data_ptr() if want_f64 else None, d32.data_ptr() if want_f32 else None, _stream()))
    return d64, d32


def mulaw_encode_t(x: torch.Tensor, mu: int = 256) -> torch.Tensor:
    _need_cuda(x)
    x = x.to(torch.float64).contiguous()
    y = torch.empty(x.shape, dtype=torch.int64, device=x.device)
    check(lib.qp_mulaw_encode(x.data_ptr(), x.numel(), mu, y.data_ptr(), _stream()))
    return y


def mulaw_decode_t(y: torch.Tensor, mu: int = 256) -> torch.Tensor:
    _need_cuda(y)
    y = y.to(torch.int64).contiguous()
    x = torch.empty(y.shape, dtype=torch.float64, device=y.device)
    check(lib.qp_mulaw_decode(y.data_ptr(), y.numel(), mu, x.data_ptr(), _stream()))
    return x


def encode_mu_law(x, mu=256):
    """qpnet.py:22-32 -- numpy in, numpy int64 out, computed on the device."""
    x = np.asarray(x, dtype=np.float64)
    return mulaw_encode_t(torch.from_numpy(np.ascontiguousarray(x)).cuda(), mu).cpu().numpy()


def decode_mu_law(y, mu=256):
    """qpnet.py:34-45 -- numpy in, numpy float64 out, computed on the device."""
    y = np.asarray(y, dtype=np.int64)
    return mulaw_decode_t(torch.from_numpy(np.ascontiguousarray(y)).cuda(), mu).cpu().numpy()


def feat_prepare(raw: torch.Tensor, n_frames: torch.Tensor, mean: torch.Tensor, scale: torch.Tensor, f0_factor: float,
                 f0_dim: int, fs: float, dense_factor: float, upsampling: int, want_f64: bool = True, want_f32: bool = False):
    """Decode front end on the device (qpnet_decode.py:163-200, 268-269): raw (B, Fmax, D) fp64 zero padded,
    n_frames (B,) -> (h (B, D, Fmax) fp32, d64 (B, Fmax*U) fp64 or None, d32 or None)."""
    _need_cuda(raw, n_frames, mean, scale)
    raw = raw.to(torch.float64).contiguous()
    B, Fmax, D = raw.shape
    n_frames = n_frames.to(torch.int32).contiguous()
    mean, scale = mean.to(torch.float64).contiguous(), scale.to(torch.float64).contiguous()
    h = torch.empty((B, D, Fmax), dtype=torch.float32, device=raw.device)
    d64 = torch.empty((B, Fmax * upsampling), dtype=torch.float64, device=raw.device) if want_f64 else None
    d32 = torch.empty((B, Fmax * upsampling), dtype=torch.float32, device=raw.device) if want_f32 else None
    check(lib.qp_feat_prepare(raw.data_ptr(), n_frames.data_ptr(), B, Fmax, D, mean.data_ptr(), scale.data_ptr(),
                              float(f0_factor), int(f0_dim), float(fs), float(dense_factor), int(upsampling), h.data_ptr(),
                              d64.data_ptr() if want_f64 else None, d32.data_ptr() if want_f32 else None, _stream()))
    return h, d64, d32


def mulaw_decode_pcm16(sym: torch.Tensor, mu: int = 256) -> torch.Tensor:
    """Decode back end (qpnet_decode.py:315-318): int32 symbols -> int16 PCM, on the device."""
    _need_cuda(sym)
    sym = sym.to(torch.int32).contiguous()
    out = torch.empty(sym.shape, dtype=torch.int16, device=sym.device)
    check(lib.qp_mulaw_decode_pcm16(sym.data_ptr(), sym.numel(), mu, out.data_ptr(), _stream()))
    return out


def cross_entropy(logits: torch.Tensor, target: torch.Tensor, want_grad: bool = True):
    """Mean softmax cross-entropy over all rows (qpnet_train.py:430,526) and its gradient."""
    _need_cuda(logits, target)
    Q = logits.shape[-1]
    lg = logits.reshape(-1, Q).contiguous().float()
    tg = target.reshape(-1).contiguous().to(torch.int64)
    rows = lg.shape[0]
    loss = torch.zeros(1, dtype=torch.float32, device=lg.device)
    dl = torch.empty_like(lg) if want_grad else None
    check(lib.qp_cross_entropy(lg.data_ptr(), tg.data_ptr(), rows, Q, 1.0 / max(rows, 1), loss.data_ptr(),
                               dl.data_ptr() if want_grad else None, _stream()))
    return loss / max(rows, 1), (dl.reshape(logits.shape) if want_grad else None)


def last_launch_count() -> int:
    return lib.qp_last_launch_count()
