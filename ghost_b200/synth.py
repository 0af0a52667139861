"""Synthetic chirp-plus-pink-noise recordings.

Per channel ``c`` the generator is seeded with ``1000 + c``; the signal is a
linear chirp from 1 Hz to 0.4*fs plus 0.5 x unit-variance pink noise.  For long
recordings the noise is shaped in independent 2**20-sample blocks.
"""
from __future__ import annotations

import numpy as np

_BLOCK = 1 << 20


def _pink(rng, n):
    white = rng.standard_normal(n)
    spec = np.fft.rfft(white)
    f = np.arange(spec.size, dtype=np.float64)
    f[0] = 1.0
    spec /= np.sqrt(f)
    spec[0] = 0.0
    noise = np.fft.irfft(spec, n=n)
    std = noise.std()
    return noise / std if std > 0 else noise


def chirp_pink(n_samples, fs, channel=0, dtype=np.float32):
    """One channel of the benchmark signal, shape (n_samples,)."""
    rng = np.random.default_rng(1000 + int(channel))
    t = np.arange(n_samples, dtype=np.float64) / fs
    dur = n_samples / fs
    f_a, f_b = 1.0, 0.4 * fs
    x = np.sin(2 * np.pi * (f_a * t + 0.5 * (f_b - f_a) / dur * t * t))
    if n_samples <= (1 << 24):
        x += 0.5 * _pink(rng, n_samples)
    else:
        for lo in range(0, n_samples, _BLOCK):
            hi = min(n_samples, lo + _BLOCK)
            x[lo:hi] += 0.5 * _pink(rng, hi - lo)
    return x.astype(dtype)


def recording(n_channels, n_samples, fs, dtype=np.float32, first_channel=0):
    """(n_channels, n_samples) array, channel-major (each channel contiguous)."""
    out = np.empty((n_channels, n_samples), dtype=dtype)
    for c in range(n_channels):
        out[c] = chirp_pink(n_samples, fs, first_channel + c, dtype)
    return out
