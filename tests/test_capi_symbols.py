"""The C-ABI library loads and exports every symbol include/polystokes_b200.h declares; without a GPU it fails loudly."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "polystokes_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ps_[a-z_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    from polystokes_b200 import _capi
    lib = ctypes.CDLL(_capi.PRODUCT_LIB)
    names = declared_symbols()
    assert set(names) == set(_capi.SYMBOLS), "ctypes binding and header disagree"
    for n in names:
        assert hasattr(lib, n), f"{n} not exported by libpolystokes_b200.so"


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from polystokes_b200 import PolyStokesSolver, PolyStokesError
    with pytest.raises(PolyStokesError) as e:
        PolyStokesSolver(8, 8, 8, 0.1, 0.01, 1000.0)
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_invalid_arguments_are_reported():
    from polystokes_b200 import _capi
    lib = _capi.load()
    h = ctypes.c_void_p()
    assert lib.ps_create(None, ctypes.byref(h)) == _capi.PS_INVALID
    assert b"null" in lib.ps_last_error()
    assert lib.ps_step(None, None, None, None) == _capi.PS_INVALID


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under polystokes_b200/ may reference it."""
    pkg = os.path.join(ROOT, "polystokes_b200")
    for dp, _, files in os.walk(pkg):
        if "build" in dp:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".hpp", ".h", ".cuh")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "ps_oracle" not in txt, f"{f} references the oracle"
