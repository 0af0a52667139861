import ctypes as C

import numpy as np

from .. import _lib

DP = C.POINTER(C.c_double)


def as_f64(a):
    """(array as float64 or complex128 C-contiguous, is_complex flag)."""
    a = np.asarray(a)
    if np.iscomplexobj(a):
        return np.ascontiguousarray(a, dtype=np.complex128), 1
    return np.ascontiguousarray(a, dtype=np.float64), 0


def ptr(a):
    return a.ctypes.data_as(DP)


def lib():
    return _lib.load()


check = _lib.check
