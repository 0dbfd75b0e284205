"""Training step around ``QPNet.forward`` (reference: ``src/bin/qpnet_train.py:517-531``) and its
data-parallel form (reference: ``nn.DataParallel``, ``qpnet_train.py:416-423``).

    loss = CE(model(x, h, d, blength)[:, -bl:].view(-1, Q), t[:, -bl:].view(-1))
    optimizer.zero_grad(); loss.backward(); optimizer.step()

Here the loss and its gradient come from one fused softmax-CE kernel (``qp_cross_entropy``),
``backward`` is the hand-written ``qp_backward`` and, with more than one rank, the replicas'
gradients are averaged by ONE all-reduce over a flat fp32 bucket -- the exchange step the reference
gets implicitly from DataParallel's ``reduce_add``.  One process per GPU (torchrun); every rank
feeds its own segment stream (the reference's DataParallel scatters one segment per GPU).

``segment_geometry`` restates the segmenter's length arithmetic (``qpnet_train.py:268-284``) so a
caller can cut segments the reference's way.
"""
from __future__ import annotations

import math

import os

import torch
import torch.distributed as dist

from . import ops


def segment_geometry(max_d: float, batch_length: int, upsampling: int, rf_causal: int, rf_fixed: int,
                     rf_adaptive: int, max_length: int = 30000):
    """(receptive_field, bl, h_frames, x_samples) for one training segment (qpnet_train.py:268-284).

    receptive_field = causal + fixed + adaptive * ceil(max d in the buffer); bl shrinks when
    R + bl would exceed ``max_length`` and so that R + bl is a whole number of frames."""
    R = rf_causal + rf_fixed + rf_adaptive * int(math.ceil(max_d))
    bl = batch_length
    if R + bl > max_length:
        bl = max_length - R
    bl -= (R + bl) % upsampling
    if bl <= 0:
        raise ValueError("receptive field leaves no room for a segment")
    h_bs = (R + bl) // upsampling
    return R, bl, h_bs, h_bs * upsampling + 1


class GradBucket:
    """Flat fp32 bucket over the gradients of ``params`` (state_dict order), averaged across ranks with
    a single all-reduce.  Parameters that never receive a gradient (the dead last ``resA_1x1``, SURVEY.md
    caveat C7) contribute zeros, so every rank reduces the same number of elements."""

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.numel = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    @property
    def world(self) -> int:
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def pack(self):
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            else:
                v.copy_(p.grad)

    def flat_in_place(self, flat):
        """True when every p.grad is a view into ``flat`` at the layout of ``qpnet.flat_layout`` (what the hand-written
        backward returns): the buffer can then be reduced where it is."""
        if flat is None or flat.dtype != torch.float32:
            return False
        from .qpnet import flat_layout
        offs, total = flat_layout(self.params)
        if flat.numel() != total:
            return False
        base = flat.data_ptr()
        return all(p.grad is not None and p.grad.is_contiguous() and p.grad.data_ptr() == base + 4 * o
                   for p, o in zip(self.params, offs))

    def allreduce_mean(self, flat=None):
        """sum over ranks, then divide: every rank ends with the mean gradient (and writes it back).  ``flat``: the
        buffer the gradients already live in (one all-reduce, no copies); otherwise they are packed into the bucket."""
        if self.flat_in_place(flat):
            w = self.world
            if w > 1:
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
                flat.div_(w)
            return flat
        self.pack()
        w = self.world
        if w > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.div_(w)
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                p.grad = v.clone()
            else:
                p.grad.copy_(v)
        return self.flat


class Trainer:
    """One SI-QPNet optimisation step per call (qpnet_train.py:517-531), data parallel when
    ``torch.distributed`` is initialised."""

    def __init__(self, model, lr: float = 1e-4, group=None):
        self.model = model
        # qpnet_train.py:426-428 (Adam, wd 0): same update rule and state_dict; on the device torch's fused multi-tensor
        # kernel updates all 216 tensors in one pass (plumbing, like the all-reduce: QPNET_FUSED_ADAM=0 for the foreach one)
        params = list(model.parameters())
        fused = params[0].is_cuda and os.environ.get("QPNET_FUSED_ADAM", "1") != "0"
        self.optimizer = torch.optim.Adam(params, lr=lr, fused=True) if fused else torch.optim.Adam(params, lr=lr)
        self.bucket = GradBucket(list(model.parameters()), group)

    def step(self, x, h, d, t, bl: int):
        """x (B, T) long, h (B, A, T/U), d (B, T) fp32, t (B, >= bl) long targets -> mean CE loss (0-dim tensor)."""
        model = self.model
        blt = torch.full((x.shape[0],), bl, dtype=torch.long, device=x.device)
        self.optimizer.zero_grad(set_to_none=True)
        logits = model(x, h, d, blt)                                      # (B, bl, Q), autograd-attached
        loss, dlogits = ops.cross_entropy(logits.detach(), t[:, -bl:])    # fused softmax-CE + gradient
        logits.backward(dlogits)
        if self.bucket.world > 1:
            self.bucket.allreduce_mean(getattr(model, "_flat_grad", None))
        self.optimizer.step()
        return loss.reshape(())
