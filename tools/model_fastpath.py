"""Numpy model of the device algorithms (design validation, NOT a product path).

Used to check, on the CPU, that the B200 algorithm -- closed-form multiplier,
half-band decimation pyramid, band-limited ("pruned") inverse FFT, overlap-save
geometry -- reproduces the oracle within the parity bars before and while the
CUDA kernels are written.  tests/test_model.py runs it against the oracle.
"""
from __future__ import annotations

import math
import numpy as np


# ---------------------------------------------------------------- multiplier
def morse_terms(gamma, beta, omega, L, rel_floor=1e-17):
    """Non-zero L-grid spectrum samples X[k] (k_first, values)."""
    f0 = np.exp((np.log(beta) - np.log(gamma)) / gamma)
    fact = omega / f0
    u = np.linspace(0, 1 - 1 / L, L)
    w = 2 * np.pi * u / fact
    K = round(L / 2)
    with np.errstate(divide="ignore", over="ignore", invalid="ignore"):
        X = 2 * np.exp(-beta * np.log(f0) + f0 ** gamma + beta * np.log(w[:K]) - w[:K] ** gamma)
    X[0] = 0.0
    keep = np.nonzero(X > rel_floor * X.max())[0]
    k0, k1 = keep[0], keep[-1] + 1
    return int(k0), X[k0:k1].copy()


def multiplier(L, k_first, X, nfft, bins):
    """Exact real zero-phase response G at omega_j = 2 pi j / nfft for j in bins,
    plus the half-sample phase for even L.  Integer-exact phase arithmetic."""
    j = np.asarray(bins, dtype=np.int64)
    k = k_first + np.arange(len(X), dtype=np.int64)
    sN = np.sin(np.pi * ((j * L) % (2 * nfft)) / nfft)            # sin(L w / 2)
    num = j[:, None] * L - k[None, :] * nfft                       # exact ints
    den = float(nfft) * float(L)
    sgn = np.where(k % 2 == 0, 1.0, -1.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        d = L * np.sin(np.pi * (num / den))
        t = np.where(num == 0, 1.0, (sgn[None, :] * sN[:, None]) / d)
    G = (t * X[None, :]).sum(axis=1)
    if L % 2 == 0:
        H = G * np.exp(-1j * np.pi * j / nfft)
    else:
        H = G.astype(complex)
    return G, H


# ---------------------------------------------------------------- generic path
def generic_cwt(x, scales, nfft=None):
    """Full-spectrum overlap-save in one chunk (the fp64 device path's math).
    scales: list of (L, k_first, X)."""
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    lmax = max(s[0] for s in scales)
    if nfft is None:
        nfft = 1 << int(math.ceil(math.log2(n + lmax - 1)))
    o = (lmax - 1 + 1) // 2
    buf = np.zeros(nfft)
    buf[o:o + n] = x
    Y = np.fft.fft(buf)
    out = np.empty((len(scales), n), dtype=complex)
    allbins = np.arange(nfft)
    for i, (L, k0, X) in enumerate(scales):
        _, H = multiplier(L, k0, X, nfft, allbins)
        out[i] = np.fft.ifft(Y * H)[o:o + n]
    return out


# ---------------------------------------------------------------- pyramid
def halfband(T=19, atten_db=140.0):
    """Zero-phase Kaiser half-band low-pass, taps t = -T..T (T odd)."""
    t = np.arange(-T, T + 1)
    beta_k = 0.1102 * (atten_db - 8.7)
    h = 0.5 * np.sinc(t / 2.0) * np.kaiser(2 * T + 1, beta_k)
    h[T] = 0.5
    h[(t % 2 == 0) & (t != 0)] = 0.0
    odd = h[t % 2 != 0]
    h[t % 2 != 0] = odd * (0.5 / odd.sum())       # exact unit DC gain
    return h


def halfband_response(h, theta):
    T = (len(h) - 1) // 2
    t = np.arange(-T, T + 1)
    return (h[None, :] * np.cos(np.outer(theta, t))).sum(axis=1)


class Pyramid:
    """x_j[i] ~ lowpass(x)[i * 2**j]; arrays carry an index origin."""

    def __init__(self, x, levels, h, dtype=np.float64):
        self.h = h
        T = (len(h) - 1) // 2
        self.data = [np.asarray(x, dtype=dtype)]
        self.base = [0]
        for _ in range(levels):
            prev, pb = self.data[-1], self.base[-1]
            lo = (pb - T) // 2                       # floor
            hi = -((-(pb + len(prev) - 1 + T)) // 2)  # ceil
            n_out = hi - lo + 1
            # pad prev so that index 2*i + t is always in range
            pad_lo = pb - (2 * lo - T)
            pad_hi = (2 * hi + T) - (pb + len(prev) - 1)
            p = np.concatenate([np.zeros(pad_lo), prev.astype(np.float64), np.zeros(pad_hi)])
            full = np.convolve(p, h[::-1], mode="valid")      # correlation, h symmetric
            self.data.append(full[::2][:n_out].astype(dtype))
            self.base.append(lo)

    def fetch(self, level, start, count):
        """x_level[start : start+count] with zeros outside the stored range."""
        arr, b = self.data[level], self.base[level]
        out = np.zeros(count, dtype=arr.dtype)
        lo = max(start, b)
        hi = min(start + count, b + len(arr))
        if hi > lo:
            out[lo - start:hi - start] = arr[lo - b:hi - b]
        return out


# ---------------------------------------------------------------- planner
MBINS = 256


def band_ok(L, k0, X, nc_full, tol):
    """Is the filter's energy outside bins [0, MBINS) of the nc_full grid below tol^2?"""
    G, _ = multiplier(L, k0, X, nc_full, np.arange(MBINS))
    e_in = float((G * G).sum())
    e_tot = nc_full / L * float((X * X).sum())
    return (1.0 - e_in / e_tot) < tol * tol, G


def plan_scale(L, k0, X, tol=3e-7, max_level=20, nc_dec=1024, max_halo_frac=0.45):
    """Pick the decimation level for one scale; level -1 means FULL (4096, two-sided)."""
    best = -1
    for lev in range(0, max_level + 1):
        D = 1 << lev
        nc_full = nc_dec * D
        if L - 1 > max_halo_frac * nc_full:
            continue
        ok, _ = band_ok(L, k0, X, nc_full, tol)
        if ok:
            best = lev
        else:
            if best >= 0:
                break
    return best


# ---------------------------------------------------------------- fast path
def fast_cwt(x, scales, tol=3e-7, nc_dec=1024, full_n=4096, dtype=np.float64, h=None,
             min_level=2, info=None):
    """Model of the fp32 device pipeline (run with dtype=np.float32 to include
    storage rounding of the pyramid and spectra)."""
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    if h is None:
        h = halfband()
    levels = [plan_scale(L, k0, X, tol, nc_dec=nc_dec) for (L, k0, X) in scales]
    # P = nc_dec*D/256 must be >= 16 for the banded kernel: level >= min_level
    levels = [lv if lv >= min_level else -1 for lv in levels]
    if info is not None:
        info["levels"] = levels
    pyr = Pyramid(x, max(max(levels), 0), h, dtype=dtype)
    out = np.empty((len(scales), n), dtype=complex)
    cdt = np.complex64 if dtype == np.float32 else np.complex128
    for lev in sorted(set(levels)):
        members = [i for i, lv in enumerate(levels) if lv == lev]
        lmax = max(scales[i][0] for i in members)
        if lev < 0:
            D, nc_full, ncd = 1, full_n, full_n
        else:
            D, nc_full, ncd = 1 << lev, nc_dec << lev, nc_dec
        align = max(D, 16)
        o = -(-((lmax - 1 + 1) // 2) // align) * align
        hop = ((nc_full - (lmax - 1) // 2 - o) // align) * align
        assert hop > 0, (lev, lmax)
        if info is not None:
            info.setdefault("classes", []).append((lev, len(members), lmax, nc_full, hop / nc_full))
        tabs = {}
        for i in members:
            L, k0, X = scales[i]
            if lev < 0:
                _, H = multiplier(L, k0, X, nc_full, np.arange(nc_full))
            else:
                _, H = multiplier(L, k0, X, nc_full, np.arange(MBINS))
                wm = 2 * np.pi * np.arange(MBINS) / nc_full
                Hd = np.ones(MBINS)
                for st in range(lev):
                    Hd *= halfband_response(h, wm * (1 << st))
                H = H / Hd
            tabs[i] = H.astype(cdt)
        q = 0
        while q * hop < n:
            t0 = q * hop - o
            chunk = pyr.fetch(max(lev, 0), t0 // D, ncd)
            Y = np.fft.fft(chunk.astype(np.float64)).astype(cdt)
            lo, hi = q * hop, min(n, (q + 1) * hop)
            for i in members:
                if lev < 0:
                    v = np.fft.ifft(Y * tabs[i])
                else:
                    Z = np.zeros(nc_full, dtype=complex)
                    Z[:MBINS] = Y[:MBINS] * tabs[i]
                    v = np.fft.ifft(Z) * D
                out[i, lo:hi] = v[lo - t0:hi - t0]
            q += 1
    return out


# ---------------------------------------------------------------- coarse grid + interpolation
def interp_taps(log2u, n_taps, oversampling):
    """Taps of the device's polyphase interpolator, from the library's host-only entry point."""
    import ctypes as C
    from ghost_b200 import _lib
    U = 1 << log2u
    taps = np.zeros((U, n_taps), dtype=np.float32)
    rc = _lib.load().gcwt_interp_taps(log2u, n_taps, float(oversampling), taps.ctypes.data_as(C.POINTER(C.c_float)))
    if rc != 0:
        raise RuntimeError("gcwt_interp_taps failed")
    return taps


def interpolate_power(power_full, log2u, taps):
    """What the amplitude / power kernels do with |W|^2: keep it on the grid n = iota * U only and
    bring it back to the full rate with the polyphase taps (output iota * U + phi =
    sum_t taps[phi][t] * p[iota + t - (T/2 - 1)]).  Returns the re-interpolated |W|^2."""
    U, T = taps.shape
    n = len(power_full)
    p = power_full[::U].astype(np.float64)
    pad = np.concatenate([np.zeros(T), p, np.zeros(T)])
    out = np.zeros(len(p) * U)
    idx = np.arange(len(p))
    for phi in range(U):
        acc = np.zeros(len(p))
        for t in range(T):
            acc += float(taps[phi, t]) * pad[T + idx + t - (T // 2 - 1)]
        out[phi::U] = acc
    return np.maximum(out[:n], 0.0)
