"""Drop-in mirror of the reference model class (``/root/reference/src/nets/qpnet.py``).

``from qpnet_b200.qpnet import QPNet, initialize, encode_mu_law, decode_mu_law`` replaces
``from qpnet import ...`` (qpnet_train.py:35-37, qpnet_decode.py:32-34).  The class keeps

* the constructor signature and every public attribute callers read
  (qpnet.py:174-237; ``receptiveCausal_field`` ... are read at qpnet_train.py:463-465),
* parameter names and shapes, so ``state_dict()`` / ``load_state_dict()`` / Adam /
  ``model.apply(initialize)`` work on the reference's checkpoints unchanged,
* ``forward(x, h, dilated_factors, blength)`` (qpnet.py:239-312) -> (B, bl, Q) logits
  attached to autograd,
* ``batch_fast_generate(x, h, n_samples_list, dilated_factors, intervals, mode,
  extra_memory)`` (qpnet.py:314-559) -> list of int64 ndarrays in finish order.

All arithmetic happens in libqpnet_b200.so (hand-written sm_100a kernels) through the C
ABI of include/qpnet_b200.h.  The submodules below only hold parameters; they are never
called.  There is no CPU path: tensors must live on a CUDA device.

Deliberate differences from the reference (SURVEY.md caveats):
* C1: every batch element gathers its past taps from ITSELF (the reference reads batch
  element 0 for all of them, qpnet.py:250; its shipped configuration is batch_size 1).
* C6: ``batch_fast_generate`` accepts a one-sample seed only (every reference caller
  passes one, qpnet_decode.py:170).
* sampling uses an inverse-CDF draw on counter-based (Philox) or caller-supplied uniforms
  instead of torch's global-RNG multinomial (qpnet.py:508-510).
"""
from __future__ import annotations

import logging
import sys

import numpy as np
import torch
from torch import nn

from . import _lib, ops
from ._lib import check, lib
from .ops import decode_mu_law, encode_mu_law  # noqa: F401  (re-exported like the reference module)


def initialize(m):
    """qpnet.py:47-58 -- Xavier-uniform Conv1d weights / zero biases; unit upsampler."""
    if isinstance(m, nn.Conv1d):
        nn.init.xavier_uniform_(m.weight)
        nn.init.constant_(m.bias, 0.0)
    if isinstance(m, nn.ConvTranspose2d):
        nn.init.constant_(m.weight, 1.0)
        nn.init.constant_(m.bias, 0.0)


class _Tap2(nn.Module):
    """Parameter holder named like CausalConv1d (qpnet.py:110-132): ``.conv`` (out, in, 2)."""

    def __init__(self, cin, cout, k):
        super().__init__()
        self.conv = nn.Conv1d(cin, cout, k)


class _CurPrev(nn.Module):
    """Parameter holder named like DilatedConv1d (qpnet.py:89-108): ``.convC`` / ``.convP``."""

    def __init__(self, cin, cout):
        super().__init__()
        self.convC = nn.Conv1d(cin, cout, 1)
        self.convP = nn.Conv1d(cin, cout, 1)


class _Up(nn.Module):
    """Parameter holder named like UpSampling (qpnet.py:134-158)."""

    def __init__(self, factor):
        super().__init__()
        self.conv = nn.ConvTranspose2d(1, 1, kernel_size=(1, factor), stride=(1, factor))


def _stream():
    return torch.cuda.current_stream().cuda_stream


def flat_layout(tensors, order=None):
    """Element offsets of ``tensors`` (indexed like ``tensors``) inside one flat buffer, each piece starting on a 16-byte
    boundary.  ``order``: the sequence in which the tensors are laid out (default: as given)."""
    offs, off = [0] * len(tensors), 0
    for i in (range(len(tensors)) if order is None else order):
        offs[i] = off
        off += (tensors[i].numel() + 3) // 4 * 4
    return offs, off


def flat_grad_views(tensors, order=None):
    offs, total = flat_layout(tensors, order)
    flat = torch.zeros(total, dtype=torch.float32, device=tensors[0].device)
    return flat, [flat[o:o + t.numel()].view(t.shape) for o, t in zip(offs, tensors)]


class _TeacherForced(torch.autograd.Function):
    """logits = stack(x, h, d; params) with the hand-written backward (qp_forward / qp_backward)."""

    @staticmethod
    def forward(ctx, model, x, h, d, bl, M, check_range, need_grad, *params):
        B, T = x.shape
        F = h.shape[2]
        flags = _lib.QP_F_SAVE if need_grad else 0
        if model.tensor_cores:
            flags |= _lib.QP_F_BF16
        arch = model._arch
        nbytes = lib.qp_forward_workspace_bytes(arch, B, T, bl, M, flags)
        if nbytes == 0:
            raise ValueError("qp_forward_workspace_bytes rejected the shapes")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        logits = torch.empty((B, bl, model.n_quantize), dtype=torch.float32, device=x.device)
        tensors = [p.detach() for p in params]
        check(lib.qp_forward(arch, _lib.ptr_array(tensors), x.data_ptr(), h.data_ptr(), d.data_ptr(), B, T, F, bl, M,
                             logits.data_ptr(), ws.data_ptr(), nbytes, flags, _stream()))
        model.last_launches = lib.qp_last_launch_count()
        if check_range:
            check(lib.qp_workspace_status(ws.data_ptr(), _stream()))   # qpnet.py:294 assert
        ctx.model, ctx.ws, ctx.nbytes, ctx.flags = model, ws, nbytes, flags
        ctx.shape = (B, T, F, bl, M)
        ctx.save_for_backward(x, h, d, *tensors)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        x, h, d, *tensors = ctx.saved_tensors
        B, T, F, bl, M = ctx.shape
        if not (ctx.flags & _lib.QP_F_SAVE):
            raise RuntimeError("forward ran without gradient tracking")
        dlogits = dlogits.contiguous().float()
        # every gradient is a view of ONE flat fp32 buffer (16-byte aligned pieces): autograd adopts the views as
        # p.grad, and the data-parallel bucket all-reduces the buffer in place instead of copying 216 tensors twice
        # The buffer is laid out in the order the backward FINISHES the tensors (head, blocks last to first, causal layer:
        # QPNet.flat_order), so a range of backward stages is a contiguous piece of it.
        model = ctx.model
        flat, grads = flat_grad_views(tensors, model.flat_order)
        model._flat_grad = flat
        sync = model.grad_sync                       # data-parallel hook (train.OverlappedReducer) or None
        ranges = sync.stage_ranges() if sync is not None else [(0, model.n_backward_stages)]
        for k, (s0, s1) in enumerate(ranges):
            check(lib.qp_backward_range(model._arch, _lib.ptr_array(tensors), x.data_ptr(), h.data_ptr(), d.data_ptr(),
                                        B, T, F, bl, M, dlogits.data_ptr(), _lib.ptr_array(grads), ctx.ws.data_ptr(),
                                        ctx.nbytes, ctx.flags, s0, s1, _stream()))
            model.last_launches += lib.qp_last_launch_count()
            if sync is not None:
                sync.range_done(k, flat)             # the gradients of stages [s0, s1) are final once the stream gets here
        ctx.ws = None
        return (None, None, None, None, None, None, None, None, *grads)


class QPNet(nn.Module):
    """QUASI-PERIODIC WAVENET -- same constructor as qpnet.py:174-178."""

    GROUP = 256     # utterances per launch of the tcgen05 cluster generators (qp_generate_f3.cu: 128, qp_generate_f3x2.cu: 256)

    def __init__(self, n_quantize=256, n_aux=39, n_resch=512, n_skipch=256,
                 dilationF_depth=4, dilationF_repeat=3, dilationA_depth=4, dilationA_repeat=1,
                 kernel_size=2, upsampling_factor=110):
        super().__init__()
        if kernel_size != 2:
            raise ValueError("only kernel_size=2 is supported (the reference's generation FIFOs assume 2 taps)")
        if upsampling_factor <= 0:
            raise ValueError("upsampling_factor must be positive")
        self.n_quantize, self.n_aux, self.n_resch, self.n_skipch = n_quantize, n_aux, n_resch, n_skipch
        self.kernel_size, self.upsampling_factor = kernel_size, upsampling_factor
        self.receptiveCausal_field = kernel_size - 1
        self.dilationF_depth, self.dilationF_repeat = dilationF_depth, dilationF_repeat
        self.dilationsF = [2 ** i for i in range(dilationF_depth)] * dilationF_repeat
        self.receptiveF_field = (kernel_size - 1) * sum(self.dilationsF)
        self.dilationA_depth, self.dilationA_repeat = dilationA_depth, dilationA_repeat
        self.dilationsA = [2 ** i for i in range(dilationA_depth)] * dilationA_repeat
        self.receptiveA_field = (kernel_size - 1) * sum(self.dilationsA)
        C_, S_, A_, Q_ = n_resch, n_skipch, n_aux, n_quantize
        # registration order == the reference's state_dict order (qpnet.py:201-235)
        self.causal = _Tap2(Q_, C_, kernel_size)
        self.upsampling = _Up(upsampling_factor)
        nF, nA = len(self.dilationsF), len(self.dilationsA)
        self.dilF_sigmoid = nn.ModuleList([_Tap2(C_, C_, kernel_size) for _ in range(nF)])
        self.dilF_tanh = nn.ModuleList([_Tap2(C_, C_, kernel_size) for _ in range(nF)])
        self.auxF_1x1_sigmoid = nn.ModuleList([nn.Conv1d(A_, C_, 1) for _ in range(nF)])
        self.auxF_1x1_tanh = nn.ModuleList([nn.Conv1d(A_, C_, 1) for _ in range(nF)])
        self.skipF_1x1 = nn.ModuleList([nn.Conv1d(C_, S_, 1) for _ in range(nF)])
        self.resF_1x1 = nn.ModuleList([nn.Conv1d(C_, C_, 1) for _ in range(nF)])
        self.dilA_sigmoid = nn.ModuleList([_CurPrev(C_, C_) for _ in range(nA)])
        self.dilA_tanh = nn.ModuleList([_CurPrev(C_, C_) for _ in range(nA)])
        self.auxA_1x1_sigmoid = nn.ModuleList([nn.Conv1d(A_, C_, 1) for _ in range(nA)])
        self.auxA_1x1_tanh = nn.ModuleList([nn.Conv1d(A_, C_, 1) for _ in range(nA)])
        self.skipA_1x1 = nn.ModuleList([nn.Conv1d(C_, S_, 1) for _ in range(nA)])
        self.resA_1x1 = nn.ModuleList([nn.Conv1d(C_, C_, 1) for _ in range(nA)])
        self.conv_post_1 = nn.Conv1d(S_, S_, 1)
        self.conv_post_2 = nn.Conv1d(S_, Q_, 1)
        self.n_ch = n_resch
        self._arch = _lib.make_arch(Q_, A_, C_, S_, upsampling_factor, self.dilationsF, self.dilationsA)
        n_expected = lib.qp_num_tensors(self._arch)
        if n_expected != len(list(self.parameters())):
            raise RuntimeError("parameter table does not match the library's layout")
        self.check_range = True     # mirror the reference's gather assert (costs one sync)
        # teacher-forced contractions on tcgen05 (bf16 operands, fp32 accumulate) when the shapes allow
        # it; False selects the exact fp32 SIMT path (tight-tolerance parity)
        self.tensor_cores = (n_resch % 64 == 0 and n_skipch % 64 == 0 and n_quantize % 32 == 0 and n_aux <= 64)
        self.last_launches = 0      # kernels launched by the most recent call (bench accounting)
        self._flat_grad = None      # flat buffer behind the parameter gradients of the most recent backward
        self.train_dtype = "bf16 (tcgen05 forward and backward GEMMs, fp32 accumulation / residual stream / gradients)"
        # widths the cluster generators are built for (qp_generate_f3.cu: tcgen05, 128 utterances per launch, any depth;
        # qp_generate_fold2.cu: mma.sync, 32 utterances, <= 16 blocks)
        self._folded_ok = (n_resch == 512 and n_skipch == 256 and n_quantize == 256 and n_aux <= 64
                           and len(self.dilationsF) >= 1 and self.dilationsF[0] == 1
                           and 4 <= len(self.dilationsF) + len(self.dilationsA) <= 64)
        self.philox_seed = 100      # qpnet_decode.py:58 default --seed
        # backward stages (qp_backward_range): 0 head, 1..L blocks last to first, L+1 causal layer + upsampler; the flat
        # gradient / parameter / Adam-moment buffers are laid out in that order
        self.n_backward_stages = nF + nA + 2
        self.grad_sync = None       # set by train.Trainer when the gradient all-reduce overlaps the backward
        self._stage_of_param = self._param_stages()
        self.flat_order = sorted(range(len(self._stage_of_param)), key=lambda i: (self._stage_of_param[i], i))

    # ------------------------------------------------------------------ helpers
    def _param_stages(self):
        """Backward stage that finishes the gradient of every parameter (state_dict order)."""
        nF, nA = len(self.dilationsF), len(self.dilationsA)
        L = nF + nA
        stages = []
        for name, _ in self.named_parameters():
            parts = name.split(".")
            if parts[0].startswith("conv_post"):
                stages.append(0)
            elif parts[0] in ("causal", "upsampling"):
                stages.append(L + 1)
            else:
                i = int(parts[1])
                layer = i if parts[0].endswith("F") or "F_" in parts[0] else nF + i
                stages.append(L - layer)
        return stages

    def flat_stage_offsets(self):
        """Element range [lo, hi) of every backward stage inside the flat gradient buffer."""
        params = list(self.parameters())
        offs, total = flat_layout(params, self.flat_order)
        lo = [total] * self.n_backward_stages
        hi = [0] * self.n_backward_stages
        for i, p in enumerate(params):
            s = self._stage_of_param[i]
            lo[s] = min(lo[s], offs[i])
            hi[s] = max(hi[s], offs[i] + (p.numel() + 3) // 4 * 4)
        return list(zip(lo, hi)), total

    def _tensors(self):
        ts = list(self.parameters())
        for t in ts:
            if not t.is_cuda:
                raise RuntimeError("QPNet parameters must be on a CUDA (sm_100a) device: call .cuda(); "
                                   "there is no CPU path")
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise RuntimeError("QPNet parameters must be contiguous float32")
        return ts

    # ------------------------------------------------------------------ qpnet.py:239-312
    def forward(self, x, h, dilated_factors, blength):
        """x (B,T) long, h (B,n_aux,T/U) float, dilated_factors (B,T) float, blength (B,) int
        -> (B, batch_length, n_quantize) logits."""
        assert torch.all(blength == blength[0])                      # qpnet.py:253
        bl = int(blength[0])
        params = self._tensors()
        ops._need_cuda(x, h, dilated_factors)
        x = x.contiguous().to(torch.int64)
        h = h.contiguous().float()
        d = dilated_factors.contiguous().float()
        M = ops.max_ceil(d)                                           # qpnet.py:255
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        return _TeacherForced.apply(self, x, h, d, bl, M, self.check_range, need_grad, *params)

    # ------------------------------------------------------------------ device-resident core
    @torch.no_grad()
    def generate_device(self, seed, h, d, n_dev, max_n, qmode=_lib.QP_MODE_SAMPLING, uniforms=None, force=None,
                        return_logits=False, check_status=True, n_host=None, utt_ids=None, pcm_out=None):
        """Generation with every input already resident in HBM: seed (B,) int64, h (B,A,F) fp32,
        d (B,F*U) fp64 (numpy flavour of the reference) or fp32 (extra_memory flavour), n_dev (B,)
        int32.  Returns (symbols (B, max_n) int32 on the device, logits or None).  ``pcm_out``: optional (B, >= max_n)
        int16 device tensor the generator fills with the decoded waveform (mu-law decode * 32768, clipped: the output
        stage of qpnet_decode.py:315-318); ``utt_ids`` (B,) int: caller-side utterance indices keying the Philox stream
        (default: the position in this call).

        The tcgen05 cluster generators run up to 256 utterances per launch (two groups of 128 above 128).  A larger batch
        of the SI default widths is dealt to launches of 256, longest first, so every launch retires at its own last step (the reference's decoder
        sorts by length for the same reason, qpnet_decode.py:257-259); ``utt_ids`` keeps every utterance on the Philox
        stream of its caller-side index, so the symbols do not depend on the grouping."""
        B = h.shape[0]
        if pcm_out is not None and (pcm_out.dtype != torch.int16 or not pcm_out.is_cuda or pcm_out.shape[0] != B
                                    or pcm_out.shape[1] < max_n or not pcm_out.is_contiguous()):
            raise ValueError("pcm_out must be a contiguous (B, >= max_n) int16 CUDA tensor")
        if B > self.GROUP and self._folded_ok:
            dev = h.device
            n_host = list(n_host) if n_host is not None else n_dev.cpu().tolist()
            order = sorted(range(B), key=lambda b: -n_host[b])
            out = torch.zeros((B, max(max_n, 1)), dtype=torch.int32, device=dev)
            logits = torch.zeros((B, max_n, self.n_quantize), dtype=torch.float32, device=dev) if return_logits else None
            launches = 0
            for c in range(0, B, self.GROUP):
                ids = order[c:c + self.GROUP]
                idx = torch.tensor(ids, dtype=torch.int64, device=dev)
                sub_max = max(n_host[b] for b in ids)
                sub_pcm = None if pcm_out is None else torch.zeros((len(ids), max(sub_max, 1)), dtype=torch.int16, device=dev)
                o, lg = self._generate_launch(seed[idx], h[idx], d[idx], n_dev[idx], sub_max, qmode,
                                              None if uniforms is None else uniforms.to(dev)[idx],
                                              None if force is None else force.to(dev)[idx], return_logits, check_status,
                                              idx.to(torch.int32) if utt_ids is None else utt_ids.to(dev)[idx].to(torch.int32),
                                              sub_pcm)
                launches += self.last_launches
                out[idx, :o.shape[1]] = o
                if pcm_out is not None:
                    pcm_out[idx, :sub_pcm.shape[1]] = sub_pcm
                if return_logits:
                    logits[idx, :sub_max] = lg
            self.last_launches = launches
            return out, logits
        return self._generate_launch(seed, h, d, n_dev, max_n, qmode, uniforms, force, return_logits, check_status, utt_ids,
                                     pcm_out)

    def _generate_launch(self, seed, h, d, n_dev, max_n, qmode, uniforms, force, return_logits, check_status, utt_ids,
                         pcm_out=None):
        """One qp_generate call (<= 256 utterances on the tcgen05 cluster kernels, any number on the generic one)."""
        params = self._tensors()
        dev = params[0].device
        ops._need_cuda(seed, h, d, n_dev)
        seed, h, d, n_dev = seed.contiguous(), h.contiguous(), d.contiguous(), n_dev.contiguous()
        B, F = h.shape[0], h.shape[2]
        if d.shape[0] != B or d.shape[1] != F * self.upsampling_factor:
            raise ValueError("dilated_factors must be (B, upsampling_factor * frames)")
        if d.dtype not in (torch.float32, torch.float64):
            raise TypeError("dilated_factors must be float32 or float64")
        M = ops.max_ceil(d)                                           # qpnet.py:347-350
        launches = ops.last_launch_count()
        out = torch.zeros((B, max(max_n, 1)), dtype=torch.int32, device=dev)
        a = _lib.QpGenerateArgs()
        a.B, a.F, a.M, a.mode, a.max_steps = B, F, M, qmode, max_n
        a.d_is_f64 = 1 if d.dtype == torch.float64 else 0
        a.seed, a.h, a.d, a.n_samples = seed.data_ptr(), h.data_ptr(), d.data_ptr(), n_dev.data_ptr()
        if uniforms is not None:
            uniforms = uniforms.to(dev).contiguous().float()
            assert uniforms.shape[0] == B and uniforms.shape[1] >= max_n
            a.uniforms, a.ld_uniforms = uniforms.data_ptr(), ops._ld(uniforms)
        a.philox_seed = int(self.philox_seed)
        if force is not None:
            force = force.to(dev).contiguous().to(torch.int32)
            assert force.shape[0] == B and force.shape[1] >= max_n
            a.force, a.ld_force = force.data_ptr(), ops._ld(force)
        a.out, a.ld_out = out.data_ptr(), ops._ld(out)
        if utt_ids is not None:
            utt_ids = utt_ids.to(dev).contiguous().to(torch.int32)
            a.utt_ids = utt_ids.data_ptr()
        if pcm_out is not None:
            a.out_pcm, a.ld_out_pcm = pcm_out.data_ptr(), ops._ld(pcm_out)
        logits = None
        if return_logits:
            logits = torch.zeros((B, max_n, self.n_quantize), dtype=torch.float32, device=dev)
            a.logits_out = logits.data_ptr()
        nbytes = lib.qp_generate_workspace_bytes(self._arch, B, M)
        if nbytes == 0:
            raise ValueError("qp_generate_workspace_bytes rejected the shapes")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        check(lib.qp_generate(self._arch, _lib.ptr_array(params), a, ws.data_ptr(), nbytes, _stream()))
        self.last_launches = launches + lib.qp_last_launch_count()
        if check_status:
            check(lib.qp_workspace_status(ws.data_ptr(), _stream()))  # sync + watchdog status
        else:
            self._last_ws = ws                                        # keep alive until the caller syncs
        return out, logits

    # ------------------------------------------------------------------ qpnet.py:314-559
    @torch.no_grad()
    def batch_fast_generate(self, x, h, n_samples_list, dilated_factors, intervals=None, mode="sampling",
                            extra_memory=False, uniforms=None, force=None, return_logits=False):
        """Batch fast generation.  Returns the list of generated int64 arrays in FINISH
        (ascending length) order and mutates ``n_samples_list`` like the reference does
        (qpnet.py:527-559): all but one entry are deleted.

        Extra keyword arguments (not in the reference): ``uniforms`` (B, >=max_n) pre-drawn
        U[0,1) numbers for the inverse-CDF sampler (default: in-kernel Philox keyed by
        ``self.philox_seed``), ``force`` (B, >=max_n) symbols fed back instead of the drawn
        ones, ``return_logits`` -> also return per-step logits (B, max_n, Q).
        """
        if mode == "sampling":
            qmode = _lib.QP_MODE_SAMPLING
        elif mode == "argmax":
            qmode = _lib.QP_MODE_ARGMAX
        else:
            logging.error("mode should be sampling or argmax")       # qpnet.py:513-515
            sys.exit(1)
        params = self._tensors()
        dev = params[0].device
        B = len(n_samples_list)
        if x.dim() != 2 or x.shape[0] != B or x.shape[1] != 1:
            raise NotImplementedError("batch_fast_generate takes a (B, 1) seed (SURVEY.md caveat C6)")
        max_n = max(n_samples_list)
        h = h.to(dev, non_blocking=True).contiguous().float()
        if extra_memory:
            d = dilated_factors.to(dev, non_blocking=True).contiguous().float()   # torch fp32 flavour (qpnet.py:615-617)
        else:                                                         # numpy fp64 flavour (qpnet.py:621-622)
            if isinstance(dilated_factors, torch.Tensor):             # (a pinned fp64 CPU tensor is accepted too)
                d = dilated_factors.to(torch.float64).to(dev, non_blocking=True).contiguous()
            else:
                d = torch.from_numpy(np.ascontiguousarray(dilated_factors, dtype=np.float64)).to(dev)
        seed = x[:, -1].to(dev).contiguous().to(torch.int64)
        n_dev = torch.tensor(list(n_samples_list), dtype=torch.int32, device=dev)
        out, logits = self.generate_device(seed, h, d, n_dev, max_n, qmode, uniforms, force, return_logits,
                                           n_host=list(n_samples_list))
        host = out.cpu().numpy().astype(np.int64)
        # ---- retirement order and caller-list mutation, exactly qpnet.py:527-557 --------
        alive = list(range(B))
        end_samples = []
        while True:
            mi = int(np.argmin(n_samples_list))
            b = alive[mi]
            end_samples.append(host[b, : n_samples_list[mi]].copy())
            if len(alive) == 1:
                break
            del alive[mi]
            del n_samples_list[mi]
        if return_logits:
            return end_samples, logits
        return end_samples
