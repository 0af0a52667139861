"""Host planner: bit-exact frequency grid / kernel lengths, spectrum terms, API errors.
No GPU needed."""
import json
import os

import numpy as np
import pytest

from ghost_b200 import ContinuousWaveletTransform, Morse
from ghost_b200.wave import morsefreq, morsehigh
from ghost_b200.utils import get_contiguous_segments
from ghost_b200.formats import standardize_input
from oracle import cwt_oracle as orc


def test_scalars_match_reference(golden_dir):
    scal = json.load(open(os.path.join(golden_dir, "morse_scalars.json")))
    for key, val in scal.items():
        g, b = (int(v) for v in key.split(","))
        assert float(morsefreq(g, b)) == val["morsefreq"]
        assert float(morsehigh(g, b)) == val["morsehigh"]


def test_grid_and_lengths_bit_exact(golden_dir):
    grids = json.load(open(os.path.join(golden_dir, "plan_grids.json")))
    for g in grids:
        cwt = ContinuousWaveletTransform(wavelet=Morse(gamma=g["gamma"], beta=g["beta"]))
        cwt.fs = g["fs"]
        cwt.wavelet.fs = g["fs"]
        f = cwt.plan_frequencies(g["n"], freq_limits=g["freq_limits"], voices_per_octave=g["vpo"])
        assert f.tolist() == g["frequencies"], g["name"]
        L = cwt.wavelet.compute_lengths(cwt._hz_to_norm_radians(f))
        assert L.tolist() == g["lengths"], g["name"]


@pytest.mark.parametrize("g,b,om,L", [(3, 20, 2.4629407752267776, 36), (3, 20, 1.9, 47),
                                      (3, 20, 0.0123, 7124), (1, 1, 0.4, 64), (9, 80, 0.3, 1013),
                                      (3, 20, 3.7e-4, 236760)])
def test_spectrum_terms_match_oracle(g, b, om, L):
    k0, X = Morse(gamma=g, beta=b).spectrum_terms(L, om)
    spec, _, _ = orc.morse_spectrum(g, b, om, L)
    full = np.zeros(L)
    full[k0:k0 + len(X)] = X
    assert np.max(np.abs(full - spec)) <= 1e-15 * spec.max()      # dropped terms are < 1e-17
    assert np.array_equal(X, spec[k0:k0 + len(X)])                # kept terms are bit-exact
    assert k0 + len(X) <= round(L / 2)


def test_freqs_keyword_uses_intended_bounds():
    cwt = ContinuousWaveletTransform()
    cwt.fs = 1250.0
    cwt.wavelet.fs = 1250.0
    want = np.geomspace(300, 1, 96)
    f = cwt.plan_frequencies(2250000, freqs=want)
    assert len(f) == 96 and np.allclose(f, np.sort(want))


def test_validation_errors_like_reference():
    x = np.zeros(2000)
    cwt = ContinuousWaveletTransform()
    with pytest.raises(TypeError):
        cwt.transform(x)                                   # fs missing
    with pytest.raises(ValueError):
        cwt.transform(x, fs=-1.0)
    with pytest.raises(ValueError):
        cwt.transform(np.zeros((2, 2000)), fs=100.0)       # two signals
    with pytest.raises(TypeError):
        cwt.transform([0.0] * 100, fs=100.0)               # not an ndarray
    with pytest.raises(ValueError):
        cwt.transform(x, fs=100.0, voices_per_octave=5)
    with pytest.raises(ValueError):
        cwt.transform(x, fs=100.0, freqs=[1, 2], freq_limits=[1, 2])
    with pytest.raises(ValueError):
        cwt.transform(x, fs=100.0, parallel="yes")
    with pytest.raises(ValueError):
        cwt.transform(x, fs=100.0, timestamps=np.zeros(7))
    with pytest.raises(ValueError):
        cwt.frequencies = [1.0]
    with pytest.raises(ValueError):
        cwt.amplitude = 3
    with pytest.raises(ValueError):
        ContinuousWaveletTransform(dtype=np.int32)
    with pytest.raises(ValueError):
        Morse(gamma=-1)
    assert cwt.amplitude is None and cwt.frequencies is None


def test_epoch_detection_matches_oracle():
    fs = 1000.0
    ts = np.arange(5000) / fs
    ts[1700:] += 0.5
    ts[4000:] += 0.0025
    a = get_contiguous_segments(ts, step=1 / fs, index=True)
    b = orc.contiguous_segments(ts, 1 / fs)
    assert a.tolist() == b.tolist() == [[0, 1700], [1700, 4000], [4000, 5000]]
    _, _, _, bounds = standardize_input(np.zeros(5000), fs=fs, timestamps=ts)
    assert bounds.tolist() == a.tolist()


class _FakeASA:
    """Duck-typed stand-in for nelpy.RegularlySampledAnalogSignalArray (nelpy is absent)."""

    def __init__(self, data, fs, lengths):
        self._data_colsig = np.asarray(data).reshape(-1, 1)
        self.n_signals = 1
        self.fs = fs
        self.lengths = np.asarray(lengths)
        self.abscissa_vals = np.arange(len(data)) / fs


def test_analog_signal_array_adapter():
    asa = _FakeASA(np.arange(10.0), 100.0, [4, 6])
    samples, fs, ts, bounds = standardize_input(asa)
    assert samples.shape == (10, 1) and fs == 100.0 and len(ts) == 10
    assert bounds.tolist() == [[0, 4], [4, 10]]            # cumulative (reference quirk Q7 fixed)
    asa.n_signals = 2
    with pytest.raises(ValueError):
        standardize_input(asa)


def test_constructor_validation_of_the_extensions():
    """dtype / output / band_tol are checked before any device is touched."""
    import pytest
    from ghost_b200 import ContinuousWaveletTransform
    with pytest.raises(ValueError):
        ContinuousWaveletTransform(dtype=np.int32)
    with pytest.raises(ValueError):
        ContinuousWaveletTransform(output="phase")
    with pytest.raises(ValueError):
        ContinuousWaveletTransform(band_tol=-1e-7)
    cwt = ContinuousWaveletTransform(dtype=np.float32, guard=False, guard_tol=1e-6, band_tol=2e-7)
    assert cwt.amplitude is None and cwt.power is None and cwt.frequencies is None


def test_time_shard_buffer_layout():
    import torch
    from ghost_b200 import sharding
    sh = sharding.TimeShard(3, 100, 7, dtype=torch.float64)
    assert sh.buf.shape == (3, 114) and sh.core.shape == (3, 100)
    sh.core.fill_(1.0)
    assert float(sh.buf[:, :7].abs().sum()) == 0.0 and float(sh.buf[:, 107:].abs().sum()) == 0.0
    assert sh.core.data_ptr() == sh.buf.data_ptr() + 7 * 8
    sharding.check_time_shards([100, 100, 50], 50)
    import pytest
    with pytest.raises(ValueError, match="halo"):
        sharding.check_time_shards([100, 100, 49], 50)
    sharding.check_time_shards([30], 50)                       # a single shard needs no halo
