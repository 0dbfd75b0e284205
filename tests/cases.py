"""Rebuild the inputs of the golden cases WITHOUT the reference (mirrors
tests/golden/make_golden.py, which produced the stored outputs from the reference)."""
import os

import numpy as np
import torch

from oracle import qpnet_oracle as orc
from qpnet_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SMALL = dict(n_resch=32, n_skipch=16)
FULL = dict()

FORWARD_CASES = {  # name: (arch, seed, bias_std, frames, bl, f0_factor, has_grads)
    "small_s1_b0": (SMALL, 1, 0.0, 12, 330, 1.0, True),
    "small_s2_b1": (SMALL, 2, 0.1, 14, 330, 1.0, True),
    "small_s3_b1_f05": (SMALL, 3, 0.1, 22, 220, 0.5, False),
    "full_s4_b1": (FULL, 4, 0.05, 12, 330, 1.0, False),
}

GENERATE_CASES = {  # name: (arch, seed, bias_std, frames_list, mode, extra_memory, f0_factor)
    "small_argmax": (SMALL, 5, 0.1, [3, 5, 4], "argmax", False, 1.0),
    "small_sampling": (SMALL, 6, 0.1, [4, 3, 5], "sampling", False, 1.0),
    "small_sampling_xm": (SMALL, 6, 0.1, [4, 3, 5], "sampling", True, 1.0),
    "small_sampling_f05": (SMALL, 7, 0.1, [6, 6], "sampling", False, 0.5),
    "small_sampling_f15": (SMALL, 8, 0.1, [5, 4], "sampling", False, 1.5),
    "full_sampling": (FULL, 9, 0.05, [2, 3], "sampling", False, 1.0),
}


def d_from_f0(f0):
    return orc.extend_time(orc.dilated_factor(f0, synth.FS, synth.DENSE_FACTOR), synth.UPSAMPLING)


def forward_inputs(name):
    kw, seed, bstd, frames, bl, fac, has_grads = FORWARD_CASES[name]
    a = orc.Arch(**kw)
    p = orc.init_params(a, seed, bstd)
    hs, f0, _ = synth.utterance(frames, seed, fac, a.A)
    T = frames * a.U
    d = torch.from_numpy(d_from_f0(f0)).float()[None, :T]
    rs = np.random.RandomState(seed)
    x = torch.from_numpy(rs.randint(0, a.Q, size=(1, T))).long()
    t = torch.from_numpy(rs.randint(0, a.Q, size=(1, bl))).long()
    h = torch.from_numpy(hs.T.copy())[None]
    return kw, a, p, x, h, d, t, bl


def generate_inputs(name):
    kw, seed, bstd, frames_list, mode, xm, fac = GENERATE_CASES[name]
    a = orc.Arch(**kw)
    p = orc.init_params(a, seed, bstd)
    B, Fmax = len(frames_list), max(frames_list)
    h = np.zeros((B, a.A, Fmax), np.float32)
    d = np.zeros((B, Fmax * a.U), np.float64)
    n_list = []
    for b, fr in enumerate(frames_list):
        hs, f0, n = synth.utterance(fr, 100 * seed + b, fac, a.A)
        h[b, :, :fr] = hs.T
        d[b, : fr * a.U] = d_from_f0(f0)
        n_list.append(n)
    x = torch.full((B, 1), a.Q // 2, dtype=torch.long)
    return kw, a, p, x, torch.from_numpy(h), d, n_list, mode, xm


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))
