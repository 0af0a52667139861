"""FFT convolution on the device (reference ghost/sigtools/convolution.py)."""
import numpy as np

from ._call import as_f64, ptr, lib, check

__all__ = ["fastconv", "fastconv_scipy", "fastconv_fftw", "fastconv_freq_scipy", "fastconv_freq_fftw"]


def _slice(res, n, m, mode):
    # convolution.py:79-87
    total = n + m - 1
    newsize = {"full": total, "same": n, "valid": n - m + 1}[mode]
    first = (total - newsize) // 2
    return res[first:first + newsize]


def fastconv(signal, kernel, *, mode=None, fft_length=None, n_threads=None, device=0):
    """Linear convolution of two 1-D arrays; ``mode`` 'full', 'same' (default) or 'valid' with
    the reference's alignment (convolution.py:16-87).  ``fft_length`` / ``n_threads`` are
    accepted for signature compatibility; the device picks its own transform size."""
    signal = np.asarray(signal)
    kernel = np.asarray(kernel)
    if signal.ndim != 1:
        raise ValueError("Signal must be 1D")
    if kernel.ndim != 1:
        raise ValueError("Kernel must be 1D")
    if mode is None:
        mode = "same"
    if mode not in ("full", "same", "valid"):
        raise ValueError("Mode must be 'full', 'same', or 'valid'")
    n, m = signal.shape[-1], kernel.shape[-1]
    if mode == "valid" and n < m:
        raise ValueError("Cannot do a 'valid' convolution because the input is shorter than the kernel")
    if fft_length is not None and fft_length < m:
        raise ValueError("FFT length must be at least the kernel size of {}".format(m))
    s, sc = as_f64(signal)
    k, kc = as_f64(kernel)
    out = np.empty(n + m - 1, dtype=np.complex128)
    check(lib().gcwt_fastconv(ptr(s), sc, n, ptr(k), kc, m, ptr(out), int(device)))
    return _slice(out, n, m, mode)


fastconv_scipy = fastconv
fastconv_fftw = fastconv


def fastconv_freq_scipy(signal_td, kernel_fd, kernel_len, *, mode=None, device=0):
    """Kernel given as frequency samples (possibly of a zero-padded kernel) of true length
    ``kernel_len`` (convolution.py:218-285): recover the taps with one inverse DFT of the
    samples' length on the device, then convolve."""
    from .fourier import dft
    signal_td = np.asarray(signal_td)
    kernel_fd = np.asarray(kernel_fd)
    if signal_td.ndim != 1:
        raise ValueError("Signal must be 1D")
    if kernel_fd.ndim != 1:
        raise ValueError("Kernel must be 1D")
    taps = dft(kernel_fd, inverse=True, device=device)[:int(kernel_len)]
    return fastconv(signal_td, taps, mode=mode, device=device)


fastconv_freq_fftw = fastconv_freq_scipy
