"""Host-side format adapters (ghost/formats): output_numpy_or_asa against the reference's behaviour
(postprocessing.py:13-65), with a stand-in nelpy module since nelpy is not installed here."""
import sys
import types

import numpy as np
import pytest

from ghost_b200.formats import output_numpy_or_asa


class _FakeRSASA:
    def __init__(self, abscissa_vals, fs, support):
        self.abscissa_vals, self.fs, self.support = abscissa_vals, fs, support


class _FakeASA:
    def __init__(self, data, *, abscissa_vals, fs, support, labels):
        self.data, self.abscissa_vals, self.fs, self.support, self.labels = data, abscissa_vals, fs, support, labels


@pytest.fixture
def fake_nelpy(monkeypatch):
    mod = types.ModuleType("nelpy")
    mod.RegularlySampledAnalogSignalArray = _FakeRSASA
    mod.AnalogSignalArray = _FakeASA
    monkeypatch.setitem(sys.modules, "nelpy", mod)
    return mod


def test_ndarray_passthrough_and_checks():
    data = np.arange(12.0).reshape(6, 2)
    assert output_numpy_or_asa(None, data) is data
    assert output_numpy_or_asa(object(), data, labels=["a", "b"]) is data      # labels ignored
    with pytest.raises(TypeError, match="Invalid output type"):
        output_numpy_or_asa(None, data, output_type="pandas")
    with pytest.raises(AttributeError):                                          # .size is read first, like the reference
        output_numpy_or_asa(None, [1, 2, 3])


def test_asa_without_nelpy(monkeypatch):
    monkeypatch.setitem(sys.modules, "nelpy", None)                              # import nelpy -> ImportError
    with pytest.raises(ModuleNotFoundError, match="nelpy"):
        output_numpy_or_asa(None, np.zeros((4, 1)), output_type="asa")


def test_asa_round_trip(fake_nelpy):
    t = np.arange(6) / 100.0
    src = _FakeRSASA(t, 100.0, support="epochs")
    data = np.arange(12.0).reshape(6, 2)
    out = output_numpy_or_asa(src, data, output_type="asa", labels=["x", "y"])
    assert isinstance(out, _FakeASA)
    assert out.data.shape == (2, 6) and np.array_equal(out.data, data.T)         # (n_signals, n_samples)
    assert out.abscissa_vals is t and out.fs == 100.0 and out.support == "epochs" and out.labels == ["x", "y"]
    with pytest.raises(TypeError, match="not a nelpy object"):
        output_numpy_or_asa(np.zeros(3), data, output_type="asa")


def test_empty_data_warns(caplog):
    with caplog.at_level("WARNING"):
        out = output_numpy_or_asa(None, np.zeros((0, 1)))
    assert out.size == 0 and "empty" in caplog.text
