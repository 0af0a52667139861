"""The C-ABI library loads and exports every symbol include/ghost_cwt.h declares.
No compute calls: there is no GPU in the build container."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from ghost_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "ghost_cwt.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gcwt_[a-z_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_lib.EXPORTS) == names
    assert lib.gcwt_version() == 100


def test_constants_match_header():
    text = open(os.path.join(ROOT, "include", "ghost_cwt.h")).read()
    defs = dict(re.findall(r"#define\s+(GCWT_[A-Z0-9_]+)\s+(-?\d+)", text))
    assert int(defs["GCWT_F32"]) == _lib.F32 and int(defs["GCWT_F64"]) == _lib.F64
    assert int(defs["GCWT_OUT_COMPLEX"]) == _lib.OUT_COMPLEX
    assert int(defs["GCWT_OUT_AMPLITUDE"]) == _lib.OUT_AMPLITUDE
    assert int(defs["GCWT_OUT_POWER"]) == _lib.OUT_POWER
    assert int(defs["GCWT_FLAG_FORCE_GENERIC"]) == _lib.FLAG_FORCE_GENERIC
    assert int(defs["GCWT_FLAG_NO_INTERP"]) == _lib.FLAG_NO_INTERP


def test_argument_errors_without_device():
    lib = _lib.load()
    handle = C.c_void_p()
    assert lib.gcwt_plan_create(C.byref(handle), None) == -1
    assert b"NULL" in lib.gcwt_last_error()
    desc = _lib.PlanDesc()
    desc.n_scales = 0
    assert lib.gcwt_plan_create(C.byref(handle), C.byref(desc)) == -1
    assert lib.gcwt_plan_destroy(None) == 0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ghost_b200 import ContinuousWaveletTransform
    with pytest.raises(_lib.GcwtError, match="no CUDA device"):
        ContinuousWaveletTransform().transform(np.zeros(4000), fs=1000.0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "ghost_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("cwt_oracle", "oracle") or f == "never", (dirpath, f)
