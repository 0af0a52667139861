"""CPU oracle for the Morse-wavelet CWT hot path of nelpy/ghost.

TEST INFRASTRUCTURE ONLY.  This module is a numpy/scipy restatement of the
reference algorithm.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and
only as the checker or as the timed CPU baseline -- never as a product path.
The product (``ghost_b200``) fails loudly when its CUDA library is missing.

Parity status: PINNED.  The reference holds no golden vectors for this path
(SURVEY.md section 4), so the pin is the running reference itself:
``oracle/gen_golden.py`` imports ``/root/reference`` (through a matplotlib stub)
in the build container, runs ``ContinuousWaveletTransform.transform``,
``Morse.__call__`` and ``fastconv_scipy`` and commits their outputs under
``tests/golden/``; ``tests/test_oracle.py`` checks every function below against
those fixtures (bit-exact for the frequency grid and kernel lengths, <= 1e-13
relative for kernels and coefficients).

Every function cites the reference ``file:line`` it restates (paths relative to
the reference root).  All arithmetic is float64 / complex128 like the reference.
"""
from __future__ import annotations

import math
from multiprocessing import cpu_count
from multiprocessing.pool import ThreadPool

import numpy as np
from scipy import fft as _sfft

__all__ = [
    "morse_peak_freq", "morse_high_freq", "freq_bounds_rad", "kernel_lengths",
    "hz_to_rad", "rad_to_hz", "clip_freq_bounds", "frequency_grid",
    "morse_spectrum", "morse_kernel", "overlap_add_same", "direct_same",
    "contiguous_segments", "cwt_complex", "cwt_amplitude", "OracleCWT",
]


# ----------------------------------------------------------------------------
# Morse wavelet scalars
# ----------------------------------------------------------------------------
def morse_peak_freq(gamma, beta):
    """Peak radian frequency of the mother wavelet.

    ghost/wave/morseutils.py:315  ``fm = exp((log(beta) - log(gamma)) / gamma)``.
    """
    return np.exp((np.log(beta) - np.log(gamma)) / gamma)


def morse_high_freq(gamma, beta, eta=0.1):
    """Highest usable peak frequency (rad/sample): first point of a 10 000-point
    scan of (1e-12, pi] at which the wavelet's value at Nyquist drops below
    ``eta`` times its peak.  Grid-quantised on purpose.

    ghost/wave/morseutils.py:607-624.
    """
    grid = np.linspace(1e-12, np.pi, 10000)
    w = morse_peak_freq(gamma, beta) * np.pi / grid
    with np.errstate(over="ignore"):
        ln_ratio = (beta / gamma) * np.log(np.exp(1) * gamma / beta) \
            + (beta * np.log(w) - w ** gamma)
    hit = np.argwhere(np.log(eta) - ln_ratio < 0).squeeze()
    return grid[np.atleast_1d(hit)[0]]


def _base_length(gamma, beta):
    # ghost/wave/morse.py:101 and :115-116 (same expression, same op order)
    w0 = morse_peak_freq(gamma, beta)
    return (2 * np.sqrt(2) * np.sqrt(gamma * beta)) / w0 * 4, w0


def freq_bounds_rad(gamma, beta, n_min, p=5):
    """[w_low, w_high] in rad/sample for a shortest segment of ``n_min``.

    ghost/wave/morse.py:93-106.
    """
    wh = morse_high_freq(gamma, beta)
    base, w0 = _base_length(gamma, beta)
    max_length = int(np.floor(n_min / p))
    max_scale = max_length / base
    return [w0 / max_scale, wh]


def kernel_lengths(gamma, beta, omegas):
    """Tap count L per scale: ceil((w0 / omega) * base_length).

    ghost/wave/morse.py:108-122.
    """
    base, w0 = _base_length(gamma, beta)
    return np.ceil((w0 / np.asarray(omegas)) * base).astype(int)


def hz_to_rad(val, fs):
    """ghost/wave/transforms.py:408-410."""
    return np.array(val) / (fs / 2.0) * np.pi


def rad_to_hz(val, fs):
    """ghost/wave/transforms.py:404-406."""
    return np.array(val) / np.pi * fs / 2.0


def clip_freq_bounds(bounds, ref):
    """Clip user [lo, hi] (Hz) into the reference bounds; the reference logs a
    warning and adjusts instead of raising.  ghost/wave/transforms.py:412-434."""
    lo, hi = bounds
    if lo < ref[0]:
        lo = ref[0]
    if hi > ref[1]:
        hi = ref[1]
    return lo, hi


def frequency_grid(fs, n_min, gamma=3, beta=20, freq_limits=None,
                   voices_per_octave=10):
    """Analysis frequencies in Hz, descending.

    ghost/wave/transforms.py:147-175 (the ``freq_limits`` / default branches).
    """
    ref = rad_to_hz(freq_bounds_rad(gamma, beta, n_min), fs)
    if freq_limits is not None:
        lim = np.sort(freq_limits)
        f_low, f_high = clip_freq_bounds([lim[0], lim[1]], ref)
    else:
        f_low, f_high = ref[0], ref[1]
    n_octaves = np.log2(f_high / f_low)
    big_j = np.floor(n_octaves * voices_per_octave)
    j = np.arange(big_j + 1)
    return f_high / 2 ** (j / voices_per_octave)


# ----------------------------------------------------------------------------
# Kernel synthesis
# ----------------------------------------------------------------------------
def morse_spectrum(gamma, beta, omega, length):
    """L-point sampled bandpass-normalised Morse spectrum X[k].

    ghost/wave/morseutils.py:115-133 (grid, log-domain evaluation, DC halving)
    and :178,188-196 (order-0 first family: unit coefficient, support only on
    bins 0 .. round(L/2)-1 with Python's banker's rounding).
    """
    f0 = morse_peak_freq(gamma, beta)
    fact = omega / f0
    w = 2 * np.pi * np.linspace(0, 1 - 1 / length, length) / fact
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        psizero = 2 * np.exp(-beta * np.log(f0) + f0 ** gamma
                             + beta * np.log(w) - w ** gamma)
    psizero[0] /= 2
    spec = np.zeros(length)
    half = round(length / 2)
    spec[:half] = psizero[:half]
    return spec, w, fact


def morse_kernel(gamma, beta, omega, length):
    """Time-domain L-tap complex kernel: inverse DFT of the centred spectrum.

    ghost/wave/morseutils.py:145-149.
    """
    spec, w, fact = morse_spectrum(gamma, beta, omega, length)
    centred = spec * np.exp(1j * w * (length + 1) / 2 * fact)
    return np.fft.ifft(centred), spec


# ----------------------------------------------------------------------------
# Convolution
# ----------------------------------------------------------------------------
def overlap_add_same(signal, kernel):
    """Zero-padded linear convolution, 'same' slice, by overlap-add with the
    reference's FFT-length rule (65536, x4 until >= len(kernel)).

    ghost/sigtools/convolution.py:41-87.
    """
    n = signal.shape[-1]
    m = kernel.shape[-1]
    total = n + m - 1
    nfft = 65536
    while nfft < m:
        nfft *= 4
    acc = np.zeros(total, dtype="<c16")
    hop = min(nfft - m + 1, n)
    kernel_f = None
    for start in range(0, n, hop):
        seg = min(hop, n - start)
        # the reference re-transforms the kernel for every block; keep that cost
        kernel_f = _sfft.fft(kernel, n=nfft)
        block = _sfft.ifft(_sfft.fft(signal[start:start + seg], n=nfft) * kernel_f)
        block = block[:seg + m - 1]
        acc[start:start + len(block)] += block
    first = (total - n) // 2
    return acc[first:first + n]


def direct_same(signal, kernel):
    """Definition of the same thing by direct summation (small cases only):
    W[n] = sum_j x[j] * psi[n - j + (L-1)//2].  SURVEY.md Appendix A."""
    full = np.convolve(np.asarray(signal, dtype=complex), kernel)
    first = (len(kernel) - 1) // 2
    return full[first:first + len(signal)]


# ----------------------------------------------------------------------------
# Epoch detection
# ----------------------------------------------------------------------------
def contiguous_segments(timestamps, step):
    """Index bounds [start, stop) of runs whose spacing is < 2*step.

    ghost/utils.py:3-42 (``index=True, inclusive=False``), including the
    reference's quirk that a break at position 0 of a 1e6 block is not seen
    (``np.any`` on indices, utils.py:29).
    """
    data = np.asarray(timestamps)
    if not np.all(data[:-1] <= data[1:]):
        data = np.sort(data)
    block = 1000000
    breaks = []
    for lo in range(0, data.size, block):
        hi = int(min(data.size, lo + block + 2))
        found = lo + np.argwhere(np.diff(data[lo:hi]) >= 2 * step)
        if np.any(found):
            breaks.extend(found)
    breaks = np.array(breaks)
    starts = np.insert(breaks + 1, 0, 0).astype(int)
    stops = np.append(breaks, len(data) - 1).astype(int)
    return np.vstack((starts, stops + 1)).T.astype(int)


# ----------------------------------------------------------------------------
# Transform
# ----------------------------------------------------------------------------
def _one_scale(x, epochs, gamma, beta, fs, freq_hz, length):
    omega = hz_to_rad(freq_hz, fs)
    kernel, _ = morse_kernel(gamma, beta, omega, int(length))
    row = np.zeros(x.shape[-1], dtype=complex)
    for a, b in epochs:
        row[a:b] = overlap_add_same(x[a:b], kernel)
    return row


def cwt_complex(x, fs, *, gamma=3, beta=20, freq_limits=None,
                voices_per_octave=10, timestamps=None, epoch_bounds=None,
                frequencies=None, parallel=False):
    """Complex coefficients (S, N), frequencies (S,), lengths (S,).

    ghost/wave/transforms.py:142-204 without the final ``np.abs`` (that is the
    oracle for complex output; the reference itself only stores magnitudes).
    """
    x = np.asarray(x).squeeze().astype(np.float64)
    x = x - np.mean(x)                       # transforms.py:142-143 (global mean)
    n = x.shape[-1]
    if epoch_bounds is None:
        if timestamps is None:
            epoch_bounds = np.array([[0, n]])
        else:
            epoch_bounds = contiguous_segments(timestamps, 1 / fs)
    epoch_bounds = np.asarray(epoch_bounds)
    n_min = int(np.min(np.diff(epoch_bounds, axis=1)))
    if frequencies is None:
        frequencies = frequency_grid(fs, n_min, gamma, beta, freq_limits,
                                     voices_per_octave)
    frequencies = np.asarray(frequencies, dtype=np.float64)
    lengths = kernel_lengths(gamma, beta, hz_to_rad(frequencies, fs))
    out = np.zeros((len(frequencies), n), dtype=complex)

    def work(i):
        out[i] = _one_scale(x, epoch_bounds, gamma, beta, fs,
                            frequencies[i], lengths[i])

    if parallel:                             # transforms.py:206-218
        pool = ThreadPool(cpu_count())
        pool.map(work, range(len(frequencies)), chunksize=1)
        pool.close()
        pool.join()
    else:
        for i in range(len(frequencies)):
            work(i)
    return out, frequencies, lengths


def cwt_amplitude(x, fs, **kw):
    """|W| as float64 (S, N): what the reference stores in ``_amplitude``
    (ghost/wave/transforms.py:204)."""
    w, f, lengths = cwt_complex(x, fs, **kw)
    return np.abs(w), f, lengths


class OracleCWT:
    """Minimal mirror of the reference object surface for timing the CPU
    baseline with the same call shape (``transform`` then ``.amplitude``)."""

    def __init__(self, gamma=3, beta=20):
        self.gamma, self.beta = gamma, beta
        self.amplitude = self.frequencies = self.time = None

    def transform(self, data, *, fs, timestamps=None, freq_limits=None,
                  voices_per_octave=None, parallel=False):
        vpo = 10 if voices_per_octave is None else voices_per_octave
        amp, f, _ = cwt_amplitude(data, fs, gamma=self.gamma, beta=self.beta,
                                  freq_limits=freq_limits, voices_per_octave=vpo,
                                  timestamps=timestamps, parallel=parallel)
        self.amplitude, self.frequencies = amp, f
        self.time = (np.arange(amp.shape[1]) / fs if timestamps is None
                     else timestamps)

    @property
    def power(self):
        return np.square(self.amplitude)
