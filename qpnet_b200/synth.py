"""Synthetic WORLD-style inputs for the QPNet hot path.

No audio, features or checkpoints ship with the reference, so every BASELINE
config is driven by synthetic tensors of the reference's shapes:

* an F0 contour per utterance (frame rate, Hz, float64), clipped to the
  speaker range of ``corpus/VCC2018/conf/pow_f0_dict.yml`` (45-450 Hz);
* a raw auxiliary matrix ``h_raw`` (F x 39): col 0 = uv flag, col 1 = F0,
  cols 2.. = N(0,1) stand-ins for mcep / coded ap
  (layout of ``src/bin/feature_extract.py:276-361``);
* the z-scored copy fed to the network (uv column untouched, as
  ``src/bin/calc_stats.py:29-37`` forces mean 0 / scale 1 there).

Only numpy is used; nothing here touches the GPU.
"""
from __future__ import annotations

import numpy as np

FS = 22050
UPSAMPLING = 110
DENSE_FACTOR = 8
N_AUX = 39
F0_MIN, F0_MAX = 45.0, 450.0


def f0_contour(n_frames: int, utt: int) -> np.ndarray:
    """Strictly positive continuous F0 contour, deterministic in ``utt``."""
    rs = np.random.RandomState(1000 + utt)
    c = rs.uniform(90.0, 260.0)
    a = rs.uniform(20.0, 0.45 * c)
    p = rs.uniform(60.0, 200.0)
    f = np.arange(n_frames, dtype=np.float64)
    f0 = c + a * np.sin(2.0 * np.pi * f / p) + 3.0 * rs.randn(n_frames)
    return np.clip(f0, F0_MIN, F0_MAX)


def raw_aux(n_frames: int, utt: int, n_aux: int = N_AUX) -> np.ndarray:
    """Unscaled aux matrix (F x n_aux, float64); col 1 is the F0 contour."""
    rs = np.random.RandomState(5000 + utt)
    h = rs.randn(n_frames, n_aux)
    uv = (rs.uniform(size=n_frames) < 0.7).astype(np.float64)
    # run-length smoothing of the voicing flag (3-frame majority)
    pad = np.pad(uv, 1, mode="edge")
    uv = ((pad[:-2] + pad[1:-1] + pad[2:]) >= 2).astype(np.float64)
    h[:, 0] = uv
    h[:, 1] = f0_contour(n_frames, utt)
    return h


def scale_aux(h_raw: np.ndarray) -> np.ndarray:
    """Fixed synthetic StandardScaler: uv untouched, F0 by (175, 60), rest as is."""
    mean = np.zeros(h_raw.shape[1])
    scale = np.ones(h_raw.shape[1])
    mean[1], scale[1] = 175.0, 60.0
    return ((h_raw - mean) / scale).astype(np.float32)


def utterance(n_frames: int, utt: int, f0_factor: float = 1.0, n_aux: int = N_AUX):
    """One decode-side utterance the way ``qpnet_decode.py:166-188`` builds it.

    Returns ``(h_scaled[F, n_aux] f32, f0[F] f64 (already scaled by f0_factor),
    n_samples)``.  The F0 factor is applied to the raw feature *before*
    z-scoring (``qpnet_decode.py:172-181``).
    """
    h = raw_aux(n_frames, utt, n_aux)
    h[:, 1] = h[:, 1] * f0_factor
    return scale_aux(h), h[:, 1].copy(), n_frames * UPSAMPLING - 1


def noise_waveform(n: int, seed: int) -> np.ndarray:
    """'Noise-shaped' stand-in waveform in [-1, 1] (white noise x 0.1, float32)."""
    rs = np.random.RandomState(9000 + seed)
    return np.clip(0.1 * rs.randn(n), -1.0, 1.0).astype(np.float32)
