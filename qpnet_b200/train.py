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
from ._lib import check, lib


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

    def __init__(self, params, group=None, order=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.order = order          # layout order of the flat gradient buffer (QPNet.flat_order)
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
        offs, total = flat_layout(self.params, self.order)
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


class FlatAdam:
    """``torch.optim.Adam(params, lr, betas, eps, weight_decay=0)`` (qpnet_train.py:426-428) as ONE hand-written kernel
    (``qp_adam_step``, qp_optim.cu) over flat buffers.

    The parameters are re-seated as views of one flat fp32 buffer (same layout as the flat gradient buffer the
    hand-written backward returns, ``qpnet.flat_layout``), so ``state_dict()`` / ``load_state_dict()`` of the model keep
    working and the step is a single element-wise pass instead of 216 tensor updates.  ``state_dict()`` /
    ``load_state_dict()`` speak torch.optim.Adam's format, so the reference's checkpoints resume here and ours resume
    there (qpnet_train.py:346-352, 481-489).  One step counter serves the whole model (torch counts per tensor that has
    seen a gradient; identical as long as every live tensor gets its gradient from the first step on, which the
    hand-written backward guarantees -- the dead last ``resA_1x1`` only ever sees zeros, and a zero gradient on zero
    moments is a zero update).  Construct it AFTER ``model.cuda()``: moving the model re-allocates its parameters and
    drops the flat buffer."""

    def __init__(self, params, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8, order=None):
        from .qpnet import flat_layout
        self.params = list(params)
        if not self.params or not all(p.is_cuda and p.dtype == torch.float32 for p in self.params):
            raise RuntimeError("FlatAdam needs float32 parameters on a CUDA (sm_100a) device; there is no CPU path")
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.offsets, self.numel = flat_layout(self.params, order)     # order: QPNet.flat_order, the layout of the flat gradients
        dev = self.params[0].device
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        for p, o in zip(self.params, self.offsets):
            view = self.flat[o:o + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self._grad = None            # packing buffer for gradients that do not already live in one flat buffer
        self.steps = 0

    def zero_grad(self, set_to_none: bool = True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def _grads_in_place(self, flat):
        if flat is None or flat.dtype != torch.float32 or flat.numel() != self.numel or not flat.is_cuda:
            return False
        base = flat.data_ptr()
        return all(p.grad is not None and p.grad.is_contiguous() and p.grad.data_ptr() == base + 4 * o
                   for p, o in zip(self.params, self.offsets))

    @torch.no_grad()
    def step(self, flat_grad=None, grad_scale: float = 1.0):
        """One Adam step.  ``flat_grad``: the buffer every ``p.grad`` is a view of (``model._flat_grad`` after the
        hand-written backward); other gradients are packed first (a missing gradient counts as zero)."""
        if not self._grads_in_place(flat_grad):
            if self._grad is None:
                self._grad = torch.zeros_like(self.flat)
            for p, o in zip(self.params, self.offsets):
                dst = self._grad[o:o + p.numel()]
                if p.grad is None:
                    dst.zero_()
                else:
                    dst.copy_(p.grad.reshape(-1))
            flat_grad = self._grad
        if self.flat.data_ptr() % 16 or any(p.data_ptr() != self.flat.data_ptr() + 4 * o for p, o in zip(self.params, self.offsets)):
            raise RuntimeError("the parameters left the flat buffer (model moved after the optimizer was built?)")
        self.steps += 1
        check(lib.qp_adam_step(self.flat.data_ptr(), flat_grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                               self.numel, self.lr, self.betas[0], self.betas[1], self.eps, self.steps, float(grad_scale),
                               torch.cuda.current_stream().cuda_stream))

    # ---- torch.optim.Adam's checkpoint format ------------------------------------------------
    def state_dict(self):
        state = {}
        if self.steps > 0:
            for i, (p, o) in enumerate(zip(self.params, self.offsets)):
                n = p.numel()
                state[i] = {"step": torch.tensor(float(self.steps)),
                            "exp_avg": self.exp_avg[o:o + n].view(p.shape).clone(),
                            "exp_avg_sq": self.exp_avg_sq[o:o + n].view(p.shape).clone()}
        group = {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": 0, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "decoupled_weight_decay": False, "params": list(range(len(self.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        group = sd["param_groups"][0]
        if len(group["params"]) != len(self.params):
            raise ValueError("optimizer state has a different number of parameters")
        if group.get("weight_decay", 0) != 0 or group.get("amsgrad", False):
            raise ValueError("only plain Adam (weight_decay 0, no amsgrad) is supported, like qpnet_train.py:426-428")
        self.lr, self.eps = float(group["lr"]), float(group["eps"])
        self.betas = (float(group["betas"][0]), float(group["betas"][1]))
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        steps = 0
        for i, (p, o) in enumerate(zip(self.params, self.offsets)):
            st = sd["state"].get(group["params"][i])
            if st is None:                      # torch keeps no state for parameters that never had a gradient (C7)
                continue
            n = p.numel()
            self.exp_avg[o:o + n].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
            steps = max(steps, int(st["step"]))
        self.steps = steps


class OverlappedReducer:
    """Gradient all-reduce overlapped with the backward (the reference's DataParallel reduces after the whole backward,
    qpnet_train.py:416-423).

    The hand-written backward runs in stage ranges (``qp_backward_range``: head, blocks last to first, causal layer) and
    returns every gradient as a view of ONE flat buffer laid out in that order, so the gradients of a finished range are
    a contiguous piece.  After each range an event is recorded on the compute stream and the piece is all-reduced
    (average) on a side stream while the next range computes; only the last, smallest piece is exposed.  Buckets shrink
    towards the end of the backward (24 MB, then 12 MB, then 6 MB) for that reason."""

    def __init__(self, model, group=None, limits_mb=(24.0, 12.0, 6.0)):
        self.model, self.group = model, group
        stage_offs, total = model.flat_stage_offsets()
        self.total = total
        self.buckets = []                       # (stage_begin, stage_end, lo, hi)
        s0, lo = 0, stage_offs[0][0]
        for s, (a, b) in enumerate(stage_offs):
            frac = b / max(total, 1)
            limit = limits_mb[0] if frac < 0.5 else (limits_mb[1] if frac < 0.85 else limits_mb[2])
            last = s == len(stage_offs) - 1
            if (b - lo) * 4 >= limit * (1 << 20) or last:
                self.buckets.append((s0, s + 1, lo, b))
                s0, lo = s + 1, b
        self.works = []
        self.comm = None

    @property
    def world(self) -> int:
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def stage_ranges(self):
        return [(b[0], b[1]) for b in self.buckets]

    def range_done(self, k, flat):
        """Called by the backward right after the launches of bucket k were enqueued."""
        w = self.world
        if w == 1:
            return
        _, _, lo, hi = self.buckets[k]
        piece = flat[lo:hi]
        nccl = flat.is_cuda
        op = dist.ReduceOp.AVG if nccl else dist.ReduceOp.SUM          # gloo (CPU tests) has no AVG
        if nccl:
            if self.comm is None:
                self.comm = torch.cuda.Stream(device=flat.device)
            ev = torch.cuda.Event()
            ev.record()
            self.comm.wait_event(ev)
            with torch.cuda.stream(self.comm):
                work = dist.all_reduce(piece, op=op, group=self.group, async_op=True)
        else:
            work = dist.all_reduce(piece, op=op, group=self.group, async_op=True)
        self.works.append((work, piece, None if nccl else w))

    def finish(self):
        """The compute stream waits for every piece; afterwards the flat buffer holds the mean gradient."""
        for work, piece, div in self.works:
            work.wait()
            if div:
                piece.div_(div)
        self.works = []


class Trainer:
    """One SI-QPNet optimisation step per call (qpnet_train.py:517-531), data parallel when
    ``torch.distributed`` is initialised."""

    def __init__(self, model, lr: float = 1e-4, group=None):
        self.model = model
        # qpnet_train.py:426-428 (Adam, wd 0): the hand-written flat-buffer kernel (FlatAdam); QPNET_TORCH_ADAM=1 selects
        # torch.optim.Adam(fused=True) on the reference-shaped tensors for A/B timing
        params = list(model.parameters())
        if os.environ.get("QPNET_TORCH_ADAM", "0") == "1":
            self.optimizer = torch.optim.Adam(params, lr=lr, fused=True)
        else:
            self.optimizer = FlatAdam(params, lr=lr, order=model.flat_order)
        self.bucket = GradBucket(list(model.parameters()), group, order=model.flat_order)
        # gradient all-reduce overlapped with the backward (QPNET_OVERLAP_ALLREDUCE=0: one all-reduce after the backward)
        self.reducer = OverlappedReducer(model, group) if os.environ.get("QPNET_OVERLAP_ALLREDUCE", "1") != "0" else None

    def step(self, x, h, d, t, bl: int):
        """x (B, T) long, h (B, A, T/U), d (B, T) fp32, t (B, >= bl) long targets -> mean CE loss (0-dim tensor)."""
        model = self.model
        blt = torch.full((x.shape[0],), bl, dtype=torch.long, device=x.device)
        self.optimizer.zero_grad(set_to_none=True)
        logits = model(x, h, d, blt)                                      # (B, bl, Q), autograd-attached
        loss, dlogits = ops.cross_entropy(logits.detach(), t[:, -bl:])    # fused softmax-CE + gradient
        overlap = self.reducer is not None and self.bucket.world > 1
        model.grad_sync = self.reducer if overlap else None
        try:
            logits.backward(dlogits)
        finally:
            model.grad_sync = None
        flat = getattr(model, "_flat_grad", None)
        if overlap:
            self.reducer.finish()
        elif self.bucket.world > 1:
            self.bucket.allreduce_mean(flat)
        if isinstance(self.optimizer, FlatAdam):
            self.optimizer.step(flat)
        else:
            self.optimizer.step()
        return loss.reshape(())


class Adapter(Trainer):
    """SD adaptation (reference: ``src/bin/qpnet_update.py:444-532``): the SI training step, started from a pretrained
    checkpoint.  ``pretrain``: a reference-format checkpoint whose ``"model"`` entry initialises the weights, iterations
    restart at 0 (qpnet_update.py:456-464); ``resume``: an adaptation checkpoint restoring model, Adam state and the
    iteration counter (445-455).  The step itself is ``Trainer.step``."""

    def __init__(self, model, pretrain=None, resume=None, lr: float = 1e-4, group=None):
        from . import checkpoint as ck
        if (pretrain is None) == (resume is None):
            raise ValueError("pass exactly one of pretrain= (start adapting) or resume= (continue adapting)")
        if pretrain is not None:
            ck.load_checkpoint(pretrain, model)           # weights only: the optimizer starts fresh
        super().__init__(model, lr=lr, group=group)       # (the parameters are re-seated into the flat buffer here)
        self.iterations = 0
        if resume is not None:
            self.iterations = ck.load_checkpoint(resume, model, self.optimizer)

    def step(self, x, h, d, t, bl: int):
        loss = super().step(x, h, d, t, bl)
        self.iterations += 1
        return loss

    def save(self, checkpoint_dir):
        from . import checkpoint as ck
        return ck.save_checkpoint(checkpoint_dir, self.model, self.optimizer, self.iterations)


@torch.no_grad()
def validation_loss(model, batches, n_quantize=None):
    """Held-out loss of one checkpoint (reference: ``src/bin/qpnet_validate.py:408-437``): the mean over the batches of
    ``CE(model(x, h, d, b)[:, -bl:], t[:, -bl:])`` -- ``float(loss / (i + 1))``, the value the reference stores in
    ``validation_result.yml`` under the checkpoint's name.  ``batches`` yields ``(x, h, t, d, b)`` like the reference's
    generators (e.g. ``segmenter.TrainSegmenter.stream``).  The forward keeps nothing for a backward (the reference's
    script leaves autograd on; the loss is the same).  Returns ``(mean, [batch losses])``."""
    Q = n_quantize or model.n_quantize
    losses = []
    for x, h, t, d, b in batches:
        assert torch.all(b == b[0])                                   # qpnet_validate.py:422
        bl = int(b[0])
        assert int(t.max()) < Q                                       # qpnet_validate.py:425
        logits = model(x, h, d, b)
        loss, _ = ops.cross_entropy(logits, t[:, -bl:], want_grad=False)
        losses.append(float(loss))
    if not losses:
        raise ValueError("no validation batches")
    return float(sum(losses) / len(losses)), losses
