"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/surfh_b200.h
declares.  No compute call is made (there is no GPU here); creating a model must fail loudly."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from surfh_b200 import _capi, build
    build.build()
    return _capi.load()


def test_every_declared_symbol_is_exported(lib):
    from surfh_b200 import _capi
    header = open(os.path.join(ROOT, "include", "surfh_b200.h")).read()
    declared = set(re.findall(r"\b(surfh_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_capi.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported"
    assert lib.surfh_abi_version() == _capi.ABI_VERSION == 5


def test_struct_layouts_match_header():
    from surfh_b200 import _capi
    # sizes implied by the header on LP64: 6 int32 + pointer + int32 (padded); int32 + int64 + 4 pointers; ...
    assert C.sizeof(_capi.ModelDesc) == 40
    assert C.sizeof(_capi.CsrDesc) == 48
    assert C.sizeof(_capi.BandDesc) == 48 + 8 + 6 * 8 + 2 * 48


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from surfh_b200 import _capi
    tpl = np.ones((2, 4))
    desc = _capi.ModelDesc(_capi.F64, 2, 8, 8, 4, 0, _capi.ptr(tpl))
    handle = C.c_void_p()
    code = lib.surfh_create(C.byref(desc), C.byref(handle))
    assert code != 0 and not handle.value
    assert b"no CUDA device" in lib.surfh_last_error(None)
    from cases import CASES
    from surfh_b200.model import spectroSigRLSCT
    cfg = CASES["mini_1band_1p"]()
    with pytest.raises(_capi.SurfhError, match="no CUDA device"):
        spectroSigRLSCT(**cfg.model_args())


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "surfh_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+(oracle|surfh_oracle)", src, re.M), f
                assert "surfh_oracle" not in src, f
