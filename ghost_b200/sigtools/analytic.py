"""Analytic signal on the device (reference ghost/sigtools/analytic.py)."""
import numpy as np

from ._call import ptr, lib, check

__all__ = ["analytic_signal", "analytic_signal_scipy", "analytic_signal_fftw"]


def analytic_signal(signal, *, fft_length=None, n_threads=None, device=0):
    """x + i*Hilbert(x) of a real 1-D signal, any length: forward DFT, zero the negative
    frequencies and double the positive ones, inverse DFT (analytic.py:22-112; same result
    as scipy.signal.hilbert)."""
    x = np.asarray(signal)
    if np.iscomplexobj(x):
        raise ValueError("signal must be real")
    x = np.ascontiguousarray(x.squeeze(), dtype=np.float64)
    if x.ndim != 1:
        raise ValueError("signal must have one non-singleton dimension")
    out = np.empty(x.size, dtype=np.complex128)
    check(lib().gcwt_analytic_signal(ptr(x), x.size, ptr(out), int(device)))
    return out


analytic_signal_scipy = analytic_signal
analytic_signal_fftw = analytic_signal
