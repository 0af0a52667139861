"""Output adapter of the drop-in: hand results back as an ndarray or wrapped in a nelpy
``AnalogSignalArray`` built around the input object.

Mirrors ``ghost/formats/postprocessing.py:13-65`` (``output_numpy_or_asa``): same arguments, same
checks in the same order, same exception types.  nelpy is optional; asking for ``'asa'`` without it
raises ``ModuleNotFoundError`` like the reference.  Host-side glue only -- no arithmetic.
"""
import logging

import numpy as np

__all__ = ["output_numpy_or_asa"]


def _nelpy():
    try:
        import nelpy
    except ImportError:
        return None
    return nelpy


def output_numpy_or_asa(obj, data, *, output_type=None, labels=None):
    """Return ``data`` itself, or an ``AnalogSignalArray`` with ``obj``'s time base around it.

    obj         : the object the data came from (ndarray or nelpy RegularlySampledAnalogSignalArray)
    data        : ndarray of shape (n_samples, n_signals)
    output_type : None (ndarray out) or 'asa'
    labels      : labels for the ASA; ignored for ndarray output
    """
    if data.size == 0:
        logging.warning("Output data is empty")
    if not isinstance(data, np.ndarray):
        raise TypeError("data must be a numpy ndarray")
    if output_type is not None and output_type != "asa":
        raise TypeError("Invalid output type {} specified".format(output_type))
    if output_type != "asa":
        return data
    nel = _nelpy()
    if nel is None:
        raise ModuleNotFoundError("You must have nelpy installed for output type {}".format(output_type))
    if not isinstance(obj, nel.RegularlySampledAnalogSignalArray):
        raise TypeError("You specified output type {} but the input object was not a nelpy object. "
                        "Cannot form an ASA around the input object".format(output_type))
    # ASAs are (n_signals, n_samples): transpose
    return nel.AnalogSignalArray(data.T, abscissa_vals=obj.abscissa_vals, fs=obj.fs,
                                 support=obj.support, labels=labels)
