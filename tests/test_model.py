"""Algorithm validation on the CPU: the numpy model of the device algorithms
(tools/model_fastpath.py: closed-form multiplier, half-band pyramid, band-limited
inverse FFT) against the oracle, at the parity bars of the north star."""
import numpy as np
import pytest

from ghost_b200 import synth, Morse
from oracle import cwt_oracle as orc
from tools import model_fastpath as mf


def _scales(gamma, beta, fs, n, **kw):
    f = orc.frequency_grid(fs, n, gamma, beta, **kw)
    om = orc.hz_to_rad(f, fs)
    L = orc.kernel_lengths(gamma, beta, om)
    m = Morse(gamma=gamma, beta=beta)
    return f, [(int(l),) + m.spectrum_terms(int(l), float(o)) for o, l in zip(om, L)]


@pytest.mark.parametrize("gamma,beta", [(3, 20), (2, 5), (9, 3), (1, 1), (6, 40)])
def test_closed_form_multiplier_fp64_bar(gamma, beta):
    fs, n = 1000.0, 6000
    x = synth.chirp_pink(n, fs, 5, np.float64)
    f, scales = _scales(gamma, beta, fs, n, voices_per_octave=4)
    W, _, _ = orc.cwt_complex(x, fs, gamma=gamma, beta=beta, frequencies=f)
    got = mf.generic_cwt(x - x.mean(), scales)
    for s in range(len(f)):
        err = np.max(np.abs(got[s] - W[s])) / np.max(np.abs(W[s]))
        assert err <= 1e-10, (gamma, beta, s, err)       # north-star fp64 bar


def test_band_limited_pipeline_fp32_bar():
    fs, n = 1000.0, 30000
    x = synth.chirp_pink(n, fs, 0, np.float32).astype(np.float64)
    f, scales = _scales(3, 20, fs, n)
    W, _, _ = orc.cwt_complex(x, fs, frequencies=f, parallel=True)
    info = {}
    got = mf.fast_cwt(x - x.mean(), scales, dtype=np.float32, info=info)
    assert max(info["levels"]) >= 4 and min(info["levels"]) == -1
    for s in range(len(f)):
        err = np.linalg.norm(np.abs(got[s]) - np.abs(W[s])) / np.linalg.norm(np.abs(W[s]))
        assert err <= 1e-5, (s, err)                     # north-star fp32 bar (rel. L2 per scale)
        assert err <= 2e-6, (s, err)                     # what the algorithm actually leaves


def test_halfband_filter_spec():
    h = mf.halfband()
    T = (len(h) - 1) // 2
    assert T == 19 and abs(h.sum() - 1.0) < 1e-15
    theta = np.linspace(0.75 * np.pi, np.pi, 2000)
    assert np.max(np.abs(mf.halfband_response(h, theta))) < 3e-7      # alias rejection
    theta = np.linspace(0, 0.25 * np.pi, 2000)
    assert np.max(np.abs(mf.halfband_response(h, theta) - 1.0)) < 3e-7


def _level_of(scale):
    return mf.plan_scale(*scale)


def test_coarse_grid_interpolation_fp32_bar():
    """Amplitude path of the device, modelled on the CPU: |W|^2 of a band-limited scale kept on the
    coarse grid U = D/2 (8 least-squares taps, over-sampling >= 4) or U = D (12 taps, >= 2.5) and
    interpolated back, against the oracle's full-rate amplitude."""
    fs, n = 1000.0, 40000
    x = synth.chirp_pink(n, fs, 1, np.float32).astype(np.float64)
    f, scales = _scales(3, 20, fs, n)
    W, _, _ = orc.cwt_complex(x, fs, frequencies=f, parallel=True)
    levels = [_level_of(s) for s in scales]
    checked = 0
    for s, lev in enumerate(levels):
        if lev < 2:
            continue
        power = np.abs(W[s]) ** 2
        for log2u, T, os_ in ((lev - 1, 8, 4.0), (lev, 12, 2.5)):
            if log2u == lev and lev > 3:                          # the wide grid is used at levels 2-3 only
                continue
            taps = mf.interp_taps(log2u, T, os_)
            amp = np.sqrt(mf.interpolate_power(power, log2u, taps))
            sl = slice(2000, n - 2000)                            # away from the zero-padded ends of this model
            err = np.linalg.norm(amp[sl] - np.abs(W[s])[sl]) / np.linalg.norm(np.abs(W[s])[sl])
            assert err <= 1e-5, (s, lev, log2u, T, err)
            assert err <= 2e-6, (s, lev, log2u, T, err)
            checked += 1
    assert checked >= 20


def test_interpolation_survives_two_tones_across_the_band():
    """Hard case for the coarse grid: a recording made of two tones on opposite flanks of one scale's
    band, so that |W|^2 of that scale is a fully modulated beat at 0.87 of its centre frequency
    (typical spectra put almost nothing there).  The tones sit where the filter is still ~3e-3 of its
    peak; further out the scale's output falls to the filter's 1e-9 leakage floor, which no
    arithmetic, fp32 or otherwise, resolves relative to such an output."""
    fs, n = 1000.0, 40000
    f, scales = _scales(3, 20, fs, n)
    levels = [_level_of(s) for s in scales]
    s = levels.index(3)                                           # first (widest) scale of level 3
    fc = f[s]
    t = np.arange(n) / fs
    x = np.sin(2 * np.pi * 0.55 * fc * t) + np.sin(2 * np.pi * 1.42 * fc * t + 0.3)
    W, _, _ = orc.cwt_complex(x, fs, frequencies=f[s:s + 1])
    power = np.abs(W[0]) ** 2
    sl = slice(4000, n - 4000)
    assert power[sl].std() > 0.3 * power[sl].mean()               # the beat is there
    for log2u, T, os_ in ((2, 8, 4.0), (3, 12, 2.5)):
        amp = np.sqrt(mf.interpolate_power(power, log2u, mf.interp_taps(log2u, T, os_)))
        err = np.linalg.norm(amp[sl] - np.abs(W[0])[sl]) / np.linalg.norm(np.abs(W[0])[sl])
        assert err <= 2e-6, (log2u, T, err)
