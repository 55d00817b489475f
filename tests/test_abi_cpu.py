"""The C-ABI library loads on a CPU-only box and exports every symbol that
include/sparrow_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from sparrowpy_b200 import build
    return build.build()


def declared_symbols():
    text = open(os.path.join(REPO, "include", "sparrow_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(spb_\w+)\s*\(", text)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    assert "spb_energy_exchange" in syms and len(syms) >= 8


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    assert lib.spb_version() >= 100


def test_layout_query_needs_no_gpu(lib_path):
    from sparrowpy_b200 import _lib
    t_pad, pad = _lib.exchange_layout(1000, 70, _lib.F64)
    assert t_pad >= 1000 and t_pad % 32 == 0 and pad >= 70 and pad % 32 == 0


def test_compute_fails_loudly_without_gpu(lib_path):
    import torch
    from sparrowpy_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.SparrowB200Error):
        _lib.call("spb_exchange_init", None, None, None, None, 0, 0, 1, 1, 0,
                  _lib.I32(0), None)
