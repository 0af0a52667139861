"""Generalised Morse wavelet: host-side planner for the device CWT.

Mirrors ``ghost.wave.Morse`` (reference ghost/wave/morse.py:12-211): same
constructor keywords, properties and ``compute_freq_bounds`` /
``compute_lengths`` results.  The float64 operation order of the reference is
kept wherever a result feeds ``ceil``/``floor``, because the
frequency grid and the tap counts must be bit-identical.

What is different: ``__call__`` does not run an L-point inverse FFT on the CPU.
The kernel is a sum of a few dozen complex exponentials; :meth:`spectrum_terms`
returns their amplitudes X[k] and the device synthesises either the kernel
(``gcwt_morse_kernel``) or, on the transform path, its transfer function in
closed form.
"""
from __future__ import annotations

import copy

import numpy as np

from .wavelet import Wavelet

__all__ = ["Morse", "morsefreq", "morsehigh"]

_TERM_FLOOR = 1e-17      # relative size below which an L-grid sample is dropped


def morsefreq(gamma, beta):
    """Peak radian frequency of the mother wavelet (morseutils.py:315)."""
    return np.exp((np.log(beta) - np.log(gamma)) / gamma)


def morsehigh(gamma, beta, eta=None):
    """High-frequency cutoff in rad/sample (morseutils.py:573-624): first point of the
    10 000-point grid on (1e-12, pi] where the wavelet at Nyquist is below ``eta``
    of its peak."""
    if eta is None:
        eta = 0.1
    if eta < 0 or eta > 1:
        raise ValueError("eta must be between 0 and 1")
    grid = np.linspace(1e-12, np.pi, 10000)
    w = morsefreq(gamma, beta) * np.pi / grid
    with np.errstate(over="ignore"):
        lnpsi = (beta / gamma) * np.log(np.exp(1) * gamma / beta) + (beta * np.log(w) - w ** gamma)
    idx = np.atleast_1d(np.argwhere(np.log(eta) - lnpsi < 0).squeeze())[0]
    return grid[idx]


class Morse(Wavelet):
    """Morse wavelet parameters (reference ghost/wave/morse.py:14-51)."""

    def __init__(self, *, fs=None, freq=None, gamma=None, beta=None):
        super().__init__()
        if fs is None:
            fs = 1
        self.fs = fs
        if freq is None:
            freq = 0.25 * self.fs
        # the reference assigns a plain attribute here and leaves norm_radian_freq
        # unset; initialise it properly
        self.frequency = freq
        if gamma is None:
            gamma = 3
        if beta is None:
            beta = 20
        self.gamma = gamma
        self.beta = beta

    # ------------------------------------------------------------------ planning
    def _base_length(self):
        w0 = morsefreq(self._gamma, self._beta)
        return (2 * np.sqrt(2) * np.sqrt(self._gamma * self._beta)) / w0 * 4, w0

    def compute_freq_bounds(self, N, *, p=None, **kwargs):
        """[low, high] usable peak frequencies in rad/sample (morse.py:93-106)."""
        if p is None:
            p = 5
        wh = morsehigh(self._gamma, self._beta, **kwargs)
        base_length, w0 = self._base_length()
        max_length = int(np.floor(N / p))
        max_scale = max_length / base_length
        wl = w0 / max_scale
        return [wl, wh]

    def compute_lengths(self, norm_radian_freqs):
        """Tap count per scale, ceil((w0 / w) * base_length) (morse.py:108-122)."""
        base_length, w0 = self._base_length()
        scale_fact = w0 / norm_radian_freqs
        return np.ceil(scale_fact * base_length).astype(int)

    def spectrum_terms(self, length, norm_radian_freq=None):
        """Non-zero samples of the L-point bandpass spectrum.

        Returns ``(k_first, X)`` with ``X[i]`` the value at L-grid bin ``k_first + i``
        (morseutils.py:115-133 for the values, :178/:194 for the support
        ``k < round(L/2)``; order-0 first-family coefficient is exactly 1).
        """
        length = int(length)
        w_s = self._norm_radian_freq if norm_radian_freq is None else norm_radian_freq
        f0 = morsefreq(self._gamma, self._beta)
        fact = w_s / f0
        half = round(length / 2)                      # Python banker's rounding, as the reference
        if length == 1:
            return 0, np.zeros(1)
        step = (1 - 1 / length) / (length - 1)        # np.linspace(0, 1-1/L, L) spacing
        kmax = min(half, 512)
        while True:
            k = np.arange(kmax)
            w = 2 * np.pi * (k * step) / fact
            with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
                X = 2 * np.exp(-self._beta * np.log(f0) + f0 ** self._gamma
                               + self._beta * np.log(w) - w ** self._gamma)
            X[0] = 0.0                                # beta*log(0) = -inf in the reference
            X[~np.isfinite(X)] = 0.0
            peak = X.max() if kmax > 0 else 0.0
            if kmax >= half or (peak > 0 and X[-1] <= _TERM_FLOOR * peak and np.argmax(X) < kmax - 1):
                break
            kmax = min(half, kmax * 4)
        if peak <= 0:
            return 0, np.zeros(1)
        keep = np.nonzero(X > _TERM_FLOOR * peak)[0]
        k0, k1 = int(keep[0]), int(keep[-1]) + 1
        return k0, np.ascontiguousarray(X[k0:k1], dtype=np.float64)

    def __call__(self, length, *, normalization=None, device=0):
        """(psi, psif): time-domain kernel of ``length`` taps and its spectrum samples
        (morse.py:53-91).  The kernel is synthesised on the GPU."""
        from .. import _lib
        import ctypes as C
        if length is None:
            length = 16384
        if length < 1:
            raise ValueError("length must at least 1 but got {}".format(length))
        if normalization is None:
            normalization = "bandpass"
        if normalization not in ("bandpass", "energy"):
            raise ValueError("normalization must be 'bandpass' or 'energy' but got {}".format(normalization))
        if normalization == "energy":
            raise NotImplementedError("only the bandpass normalisation is on the CWT path")
        length = int(length)
        k0, X = self.spectrum_terms(length)
        psif = np.zeros(length)
        psif[k0:k0 + len(X)] = X
        out = np.empty(length, dtype=np.complex128)
        lib = _lib.load()
        _lib.check(lib.gcwt_morse_kernel(length, k0, len(X), X.ctypes.data_as(C.POINTER(C.c_double)),
                                         out.ctypes.data_as(C.POINTER(C.c_double)), int(device)))
        return out, psif

    def copy(self):
        return copy.deepcopy(self)

    def _norm_radians_to_hz(self, val):
        return val / np.pi * self._fs / 2

    def _hz_to_norm_radians(self, val):
        return val / (self._fs / 2) * np.pi

    # ------------------------------------------------------------------ properties
    @property
    def fs(self):
        return self._fs

    @fs.setter
    def fs(self, val):
        if not val > 0:
            raise ValueError("fs must be positive but got {}".format(val))
        self._fs = val

    @property
    def frequency(self):
        return self._freq

    @frequency.setter
    def frequency(self, val):
        if not (val > 0 and val <= self._fs / 2):
            raise ValueError("The frequency must be between 0 and the Nyquist frequency {} Hz"
                             " but got {}".format(self._fs / 2, val))
        self._freq = val
        self._norm_radian_freq = self._hz_to_norm_radians(val)

    @property
    def norm_radian_freq(self):
        return self._norm_radian_freq

    @norm_radian_freq.setter
    def norm_radian_freq(self, val):
        if not (val > 0 and val <= np.pi):
            raise ValueError("The normalized radian frequency must be between 0 and the Nyquist"
                             " frequency pi but got {}".format(val))
        self._norm_radian_freq = val
        self._freq = self._norm_radians_to_hz(val)

    @property
    def gamma(self):
        return self._gamma

    @gamma.setter
    def gamma(self, val):
        if not val > 0:
            raise ValueError("gamma must be positive")
        self._gamma = val

    @property
    def beta(self):
        return self._beta

    @beta.setter
    def beta(self, val):
        if not val > 0:
            raise ValueError("beta must be positive")
        self._beta = val

    @property
    def time_bandwidth(self):
        return self._gamma * self._beta
