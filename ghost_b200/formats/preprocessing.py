"""Input adapter: ndarray or nelpy-style analog signal -> samples, fs, timestamps, epochs.

Same decisions as the reference decorator ``standardize_asa``
(ghost/formats/preprocessing.py:54-187) for the one call site on the CWT path
(``transform``: ``n_signals=1, class_method=True, abscissa_vals='timestamps'``), written
as a plain function.  nelpy is not required: any object exposing ``n_signals``, ``fs``,
``abscissa_vals``, ``lengths`` and ``_data_colsig`` (or ``data``) is taken as a regularly
sampled analog signal array.
"""
from __future__ import annotations

import logging

import numpy as np

from ..utils import get_contiguous_segments

__all__ = ["standardize_input", "is_analog_signal_array"]

_log = logging.getLogger("ghost")


def is_analog_signal_array(obj):
    try:
        import nelpy as nel
        if isinstance(obj, nel.RegularlySampledAnalogSignalArray):
            return True
    except Exception:
        pass
    return all(hasattr(obj, a) for a in ("n_signals", "fs", "abscissa_vals", "lengths")) \
        and (hasattr(obj, "_data_colsig") or hasattr(obj, "data"))


def standardize_input(data, *, fs=None, timestamps=None, n_signals=1, func_name="transform"):
    """Return ``(samples, fs, timestamps, epoch_bounds)``.

    ``samples`` has shape ``(n_samples, n_signals)`` (column per signal, like the
    reference); ``epoch_bounds`` is an ``(E, 2)`` int array of ``[start, stop)`` indices.
    """
    if is_analog_signal_array(data):                       # preprocessing.py:78-114
        if n_signals is not None and data.n_signals != n_signals:
            raise ValueError("Input object 'data'.n_signals=={}, but expected {}".format(
                data.n_signals, n_signals))
        if fs is not None:
            _log.warning("'fs' was passed in, but will be overwritten by the object's 'fs' attribute")
        if timestamps is not None:
            _log.warning("'timestamps' was passed in, but will be overwritten by the object's"
                         " 'abscissa_vals' attribute")
        if hasattr(data, "_data_colsig"):
            samples = np.asarray(data._data_colsig)
        else:
            samples = np.asarray(data.data).T
        # cumulative bounds; the reference builds them from raw lengths, which is only
        # right for a single epoch
        edges = np.concatenate(([0], np.cumsum(np.asarray(data.lengths, dtype=np.int64))))
        bounds = np.vstack((edges[:-1], edges[1:])).T.astype(int)
        return samples, data.fs, np.asarray(data.abscissa_vals), bounds

    if not isinstance(data, np.ndarray):                   # preprocessing.py:120-122
        raise TypeError("Input was not a nelpy.RegularlySampledAnalogSignalArray so expected a"
                        " numpy ndarray but got {}".format(type(data)))
    samples = np.atleast_1d(data.squeeze())
    if samples.ndim == 1:
        samples = samples.reshape((-1, 1))
    if n_signals is not None and samples.shape[-1] != n_signals:   # :127-130
        raise ValueError("Expected {} number of signals but got {}".format(n_signals, samples.shape[0]))
    if fs is None:                                          # :132-136
        raise TypeError("{}() missing 1 required keyword argument: 'fs'".format(func_name))
    n = samples.shape[0]
    if timestamps is None:                                  # :138-147
        ts = np.arange(n, dtype=np.float64) / fs
        bounds = np.array([[0, n]], dtype=int)
    else:
        if not isinstance(timestamps, np.ndarray):
            raise TypeError("Expected 'timestamps' to be a numpy.ndarray but got {}".format(type(timestamps)))
        if timestamps.ndim != 1:
            raise ValueError("'timestamps' should have at most one non-singleton dimension")
        if timestamps.shape[0] != n:
            raise ValueError("The argument 'timestamps' has {} sample points, but the data 'data'"
                             " has {}".format(timestamps.shape[0], n))
        ts = timestamps
        bounds = get_contiguous_segments(ts, step=1 / fs, assume_sorted=False, index=True,
                                         inclusive=False)
    return samples, fs, ts, bounds
