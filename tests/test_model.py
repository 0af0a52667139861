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
