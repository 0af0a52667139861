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
    out = np.zeros(n + m - 1, dtype=np.complex128)
    # one device call holds at most 2^26 points of signal + kernel: longer signals are convolved block-wise
    # and overlap-added here, like the reference's own loop (convolution.py:63-77) -- any length works
    limit = 1 << 26
    if m >= limit // 2:
        raise ValueError("Kernel of {} taps is too long for the device convolution (limit {})".format(m, limit // 2))
    block = n if n + m - 1 <= limit else limit - m + 1
    for start in range(0, n, block):
        length = min(block, n - start)
        seg = np.ascontiguousarray(s[start:start + length])
        part = np.empty(length + m - 1, dtype=np.complex128)
        check(lib().gcwt_fastconv(ptr(seg), sc, length, ptr(k), kc, m, ptr(part), int(device)))
        out[start:start + length + m - 1] += part
    return _slice(out, n, m, mode)


fastconv_scipy = fastconv
fastconv_fftw = fastconv


def fastconv_freq_scipy(signal_td, kernel_fd, kernel_len, *, mode=None, n_threads=None, device=0):
    """Convolution with the kernel given as frequency samples ``kernel_fd`` (convolution.py:218-285),
    block by block exactly as the reference does it: every block of ``len(kernel_fd) - kernel_len + 1``
    signal samples is zero-padded to ``len(kernel_fd)``, transformed, multiplied by ``kernel_fd`` as it
    is, transformed back, and its first ``length + kernel_len - 1`` outputs are overlap-added.  This is a
    circular convolution with the WHOLE inverse DFT of ``kernel_fd``: when those taps do not vanish beyond
    ``kernel_len`` the result keeps their wrap-around, like the reference's.  The DFTs (any length,
    Bluestein for non powers of two) run on the device; ``n_threads`` is accepted for compatibility with
    ``fastconv_freq_fftw``."""
    from .fourier import dft
    signal_td = np.asarray(signal_td)
    kernel_fd = np.asarray(kernel_fd)
    if signal_td.ndim != 1:
        raise ValueError("Signal must be 1D")
    if kernel_fd.ndim != 1:
        raise ValueError("Kernel must be 1D")
    if mode is None:
        mode = "same"
    if mode not in ("full", "same", "valid"):
        raise ValueError("Mode must be 'full', 'same', or 'valid'")
    n, m, nfd = signal_td.shape[-1], int(kernel_len), kernel_fd.shape[-1]
    if mode == "valid" and n < m:
        raise ValueError("Cannot do a 'valid' convolution because the input is shorter than the kernel")
    res = np.zeros(n + m - 1, dtype=np.complex128)
    chunk = min(nfd - m + 1, n)
    if chunk < 1:
        raise ValueError("kernel_fd must hold at least kernel_len frequency samples")
    kfd = np.asarray(kernel_fd, dtype=np.complex128)
    for start in range(0, n, chunk):
        length = min(chunk, n - start)
        padded = np.zeros(nfd, dtype=np.complex128)
        padded[:length] = signal_td[start:start + length]
        conv = dft(dft(padded, device=device) * kfd, inverse=True, device=device)[:length + m - 1]
        res[start:start + len(conv)] += conv
    return _slice(res, n, m, mode)


fastconv_freq_fftw = fastconv_freq_scipy
