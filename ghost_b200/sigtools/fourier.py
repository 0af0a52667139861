"""Discrete Fourier transforms of any length on the device (reference ghost/sigtools/fourier.py)."""
import numpy as np

from ._call import ptr, lib, check

__all__ = ["chirpz_dft", "dft"]


def dft(x, *, inverse=False, device=0):
    """DFT (or inverse DFT, scaled by 1/N) of a 1-D array of any length."""
    x = np.asarray(x)
    if x.ndim != 1:
        raise ValueError("Data must be 1-dimensional")
    xc = np.ascontiguousarray(x, dtype=np.complex128)
    out = np.empty_like(xc)
    check(lib().gcwt_dft(ptr(xc), xc.size, 1 if inverse else -1, ptr(out), int(device)))
    return out / xc.size if inverse else out


def chirpz_dft(x, *, device=0):
    """DFT through the chirp-z (Bluestein) identity, fast for prime lengths
    (reference fourier.py:9-48); powers of two skip the chirps."""
    return dft(x, device=device)
