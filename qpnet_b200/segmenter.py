"""Device-side streaming training segmenter -- the step in front of ``QPNet.forward`` in training.

Mirrors ``train_generator`` (qpnet_train.py:200-335) without its file readers: utterances arrive as arrays (int16
waveform + fp64 ``(frames, D)`` WORLD-style features, what ``wavfile.read`` / ``read_hdf5`` return) and everything after
that lives on the GPU: sample scaling and mu-law encoding (qpnet.py:22-32), z-scoring + transposition of the features
(qpnet_train.py:433-440, the decode front end's kernel), F0 -> per-sample dilated factors (147-179, utils.py:216-235), the
growing buffers (260-262), the receptive field from the largest buffered factor (181-199, 264-268) and the segment cuts
(269-316).  The host keeps only the length arithmetic.  Batches come out as the tuple the reference yields:
``(x (B, T) long, h (B, D, T/U) float, t (B, T) long, d (B, T) float, b (B,) long)``.

Differences: every element-wise transform is applied once per utterance when it is pushed instead of once per segment
(segments overlap by the receptive field, so the reference recomputes those samples); the mu-law symbols are computed in
fp64 from the float32 samples, which equals the reference's float32 arithmetic except on rounding ties
(tests/test_gpu_parity.py::test_train_segmenter_vs_reference_goldens reports the match rate).
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def validated_lengths(n_x: int, n_h: int, U: int):
    """qpnet_train.py:119-145 as length arithmetic: samples / frames kept of a (waveform, features) pair."""
    if n_x > n_h * U:
        n_x = n_h * U
    if n_x < n_h * U:
        n_h -= (n_h * U - n_x) // U + 1
        n_x = n_h * U
    return n_x, n_h


def segment_lengths(rf: int, batch_length: int, max_length: int, U: int):
    """qpnet_train.py:268-284: (segment length bl, frames h_bs, samples x_bs) for the current receptive field."""
    bl = batch_length - max(rf + batch_length - max_length, 0)
    bl -= (rf + bl) % U
    h_bs = (rf + bl) // U
    return bl, h_bs, h_bs * U + 1


class TrainSegmenter:
    def __init__(self, rf_causal: int, rf_fixed: int, rf_adaptive: int, mean, scale, fs: float, dense_factor: float = 8,
                 batch_length: int = 20000, batch_size: int = 1, max_length: int = 30000, f0_threshold: float = 0,
                 upsampling_factor: int = 110, n_quantize: int = 256, device="cuda"):
        self.rf = (int(rf_causal), int(rf_fixed), int(rf_adaptive))
        self.dev = torch.device(device)
        self.mean = torch.as_tensor(np.asarray(mean, np.float64), device=self.dev)
        self.scale = torch.as_tensor(np.asarray(scale, np.float64), device=self.dev)
        self.fs, self.dense, self.U, self.Q = float(fs), float(dense_factor), int(upsampling_factor), int(n_quantize)
        self.batch_length, self.batch_size, self.max_length = int(batch_length), int(batch_size), int(max_length)
        self.f0_threshold = float(f0_threshold)
        D = self.mean.numel()
        self.sym = torch.empty(0, dtype=torch.int64, device=self.dev)          # x_buffer, already mu-law symbols
        self.hz = torch.empty((D, 0), dtype=torch.float32, device=self.dev)    # h_buffer, already z-scored, (D, frames)
        self.d64 = torch.empty(0, dtype=torch.float64, device=self.dev)        # d_buffer (the receptive field reads fp64)
        self.d32 = torch.empty(0, dtype=torch.float32, device=self.dev)
        self._batch = [[], [], [], [], []]

    def push(self, wav, raw):
        """Append one utterance (int16 waveform, fp64 (frames, D) features) and return the batches it completes."""
        wav = torch.as_tensor(np.ascontiguousarray(wav)).to(self.dev)
        raw = torch.as_tensor(np.ascontiguousarray(raw, dtype=np.float64)).to(self.dev)
        n_x, n_h = validated_lengths(wav.numel(), raw.shape[0], self.U)
        if n_h <= 0:
            return []
        x = wav[:n_x].to(torch.float32) / 32768                                  # qpnet_train.py:249
        raw = raw[:n_h].contiguous()
        sym = ops.mulaw_encode_t(x.to(torch.float64), self.Q)
        frames = torch.tensor([n_h], dtype=torch.int32, device=self.dev)
        hz, _, _ = ops.feat_prepare(raw[None], frames, self.mean, self.scale, 1.0, 1, self.fs, self.dense, self.U,
                                    want_f64=False, want_f32=True)
        d64, d32 = ops.f0_to_dilated(raw[None, :, 1], self.fs, self.dense, self.U, f0_floor=self.f0_threshold)
        self.sym = torch.cat([self.sym, sym])
        self.hz = torch.cat([self.hz, hz[0]], dim=1)
        self.d64 = torch.cat([self.d64, d64[0]])
        self.d32 = torch.cat([self.d32, d32[0]])
        return self._cut()

    def _cut(self):
        rfC, rfF, rfA = self.rf
        U = self.U
        rf = rfF + rfA * ops.max_ceil(self.d64) + rfC                            # 181-199 on the buffered factors
        bl, h_bs, x_bs = segment_lengths(rf, self.batch_length, self.max_length, U)   # 268-284
        out = []
        bx, bh, bt, bd, bb = self._batch
        want = self.batch_size - len(bx)
        while self.hz.shape[1] > want * h_bs and self.sym.numel() > want * x_bs:
            bx.append(self.sym[: x_bs - 1]); bt.append(self.sym[1:x_bs])
            bh.append(self.hz[:, :h_bs]); bd.append(self.d32[: x_bs - 1]); bb.append(bl)
            want -= 1
            h_ss = bl // U
            self.sym, self.hz = self.sym[h_ss * U:], self.hz[:, h_ss:]
            self.d64, self.d32 = self.d64[h_ss * U:], self.d32[h_ss * U:]
            if len(bx) == self.batch_size:
                out.append((torch.stack(bx), torch.stack(bh), torch.stack(bt), torch.stack(bd),
                            torch.tensor(bb, dtype=torch.long, device=self.dev)))
                bx.clear(); bh.clear(); bt.clear(); bd.clear(); bb.clear()
                want = self.batch_size
        return out

    def stream(self, utterances, shuffle: bool = False):
        """Endless batch stream over ``utterances`` like train_generator (re-shuffled every pass when asked)."""
        order = np.random.permutation(len(utterances)) if shuffle else np.arange(len(utterances))
        while True:
            for i in order:
                for batch in self.push(*utterances[i]):
                    yield batch
            if shuffle:
                order = np.random.permutation(len(utterances))
