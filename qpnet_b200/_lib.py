"""ctypes binding of libqpnet_b200.so (the C ABI declared in include/qpnet_b200.h).

There is no fallback: if the library has not been built, importing this module raises.
Build it with ``python -m qpnet_b200.build`` (or ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libqpnet_b200.so")
QP_MAX_LAYERS = 64

QP_OK, QP_EINVAL, QP_EARCH, QP_ECUDA, QP_EWORKSPACE, QP_ERANGE, QP_ETIMEOUT = 0, -1, -2, -3, -4, -5, -6
QP_F_SAVE, QP_F_BF16 = 1, 2
QP_MODE_SAMPLING, QP_MODE_ARGMAX = 0, 1


class QpArch(C.Structure):
    _fields_ = [("n_quantize", C.c_int32), ("n_aux", C.c_int32), ("n_resch", C.c_int32),
                ("n_skipch", C.c_int32), ("upsampling", C.c_int32), ("n_fixed", C.c_int32),
                ("n_adaptive", C.c_int32), ("dil_fixed", C.c_int32 * QP_MAX_LAYERS),
                ("dil_adaptive", C.c_int32 * QP_MAX_LAYERS)]


class QpGenerateArgs(C.Structure):
    _fields_ = [("B", C.c_int32), ("F", C.c_int32), ("M", C.c_int32), ("mode", C.c_int32),
                ("max_steps", C.c_int32), ("d_is_f64", C.c_int32),
                ("seed", C.c_void_p), ("h", C.c_void_p), ("d", C.c_void_p), ("n_samples", C.c_void_p),
                ("uniforms", C.c_void_p), ("ld_uniforms", C.c_int64), ("philox_seed", C.c_uint64),
                ("force", C.c_void_p), ("ld_force", C.c_int64),
                ("out", C.c_void_p), ("ld_out", C.c_int64), ("logits_out", C.c_void_p),
                ("utt_ids", C.c_void_p), ("out_pcm", C.c_void_p), ("ld_out_pcm", C.c_int64)]


# name -> (restype, argtypes): every symbol include/qpnet_b200.h declares
_P, _I32, _I64, _U32, _F32, _F64, _SZ = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_float, C.c_double, C.c_size_t
SIGNATURES = {
    "qp_abi_version": (C.c_int, []),
    "qp_last_error": (C.c_char_p, []),
    "qp_num_tensors": (C.c_int, [C.POINTER(QpArch)]),
    "qp_device_ok": (C.c_int, []),
    "qp_mulaw_encode": (C.c_int, [_P, _I64, _I32, _P, _P]),
    "qp_mulaw_decode": (C.c_int, [_P, _I64, _I32, _P, _P]),
    "qp_f0_to_dilated": (C.c_int, [_P, _I32, _I32, _F64, _F64, _I32, _F64, _P, _P, _P]),
    "qp_feat_prepare": (C.c_int, [_P, _P, _I32, _I32, _I32, _P, _P, _F64, _I32, _F64, _F64, _I32, _P, _P, _P, _P]),
    "qp_mulaw_decode_pcm16": (C.c_int, [_P, _I64, _I32, _P, _P]),
    "qp_max_ceil_f32": (C.c_int, [_P, _I64, _P, _P]),
    "qp_max_ceil_f64": (C.c_int, [_P, _I64, _P, _P]),
    "qp_index_tf_f32": (C.c_int, [_P, _I32, _I32, _I64, _I32, _P, _P]),
    "qp_index_tf_f64": (C.c_int, [_P, _I32, _I32, _I64, _I32, _P, _P]),
    "qp_index_gen_f32": (C.c_int, [_P, _I32, _I32, _I64, _I32, _P, _P]),
    "qp_index_gen_f64": (C.c_int, [_P, _I32, _I32, _I64, _I32, _P, _P]),
    "qp_forward_workspace_bytes": (_SZ, [C.POINTER(QpArch), _I32, _I32, _I32, _I32, _U32]),
    "qp_forward": (C.c_int, [C.POINTER(QpArch), C.POINTER(_P), _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _P, _P, _SZ,
                             _U32, _P]),
    "qp_backward": (C.c_int, [C.POINTER(QpArch), C.POINTER(_P), _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _P,
                              C.POINTER(_P), _P, _SZ, _U32, _P]),
    "qp_backward_range": (C.c_int, [C.POINTER(QpArch), C.POINTER(_P), _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _P,
                                    C.POINTER(_P), _P, _SZ, _U32, _I32, _I32, _P]),
    "qp_cross_entropy": (C.c_int, [_P, _P, _I64, _I32, _F32, _P, _P, _P]),
    "qp_adam_step": (C.c_int, [_P, _P, _P, _P, _I64, _F64, _F64, _F64, _F64, _I32, _F64, _P]),
    "qp_generate_workspace_bytes": (_SZ, [C.POINTER(QpArch), _I32, _I32]),
    "qp_generate": (C.c_int, [C.POINTER(QpArch), C.POINTER(_P), C.POINTER(QpGenerateArgs), _P, _SZ, _P]),
    "qp_workspace_status": (C.c_int, [_P, _P]),
    "qp_last_launch_count": (C.c_int, []),
    "qp_debug_tma_segments": (C.c_int64, []),
}


class QpError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libqpnet_b200 error {code}: {msg}")
        self.code = code


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA library has not been built (python -m qpnet_b200.build). "
            "qpnet_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header / library mismatch
        fn.restype, fn.argtypes = res, args
    if lib.qp_abi_version() != 3:
        raise ImportError("libqpnet_b200.so ABI version mismatch")
    return lib


lib = _load()


def check(rc: int) -> int:
    """Translate a negative status into the Python exception the reference would raise."""
    if rc >= 0:
        return rc
    msg = lib.qp_last_error().decode("utf-8", "replace")
    if rc == QP_EINVAL:
        raise ValueError(msg)
    if rc == QP_ERANGE:
        raise AssertionError(msg)          # qpnet.py:294 is a Python assert
    raise QpError(rc, msg)


def make_arch(n_quantize, n_aux, n_resch, n_skipch, upsampling, dil_fixed, dil_adaptive) -> QpArch:
    a = QpArch()
    a.n_quantize, a.n_aux, a.n_resch, a.n_skipch, a.upsampling = n_quantize, n_aux, n_resch, n_skipch, upsampling
    a.n_fixed, a.n_adaptive = len(dil_fixed), len(dil_adaptive)
    if a.n_fixed > QP_MAX_LAYERS or a.n_adaptive > QP_MAX_LAYERS:
        raise ValueError("too many residual blocks")
    for i, d in enumerate(dil_fixed):
        a.dil_fixed[i] = d
    for i, d in enumerate(dil_adaptive):
        a.dil_adaptive[i] = d
    return a


def ptr_array(tensors):
    """Host array of device pointers (state_dict order) from torch tensors."""
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr
