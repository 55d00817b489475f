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


def _raw(lib_path):
    lib = ctypes.CDLL(lib_path)
    lib.spb_last_error.restype = ctypes.c_char_p
    return lib


I64, INT, PTR = ctypes.c_int64, ctypes.c_int, ctypes.c_void_p


def test_argument_errors_are_reported_not_executed(lib_path):
    """Every entry point validates its arguments before touching the device: a bad call
    returns a negative code and a message through spb_last_error (no exception, no
    abort) -- the error behaviour a binder in another language relies on."""
    lib = _raw(lib_path)
    buf = (ctypes.c_double * 64)()
    p = ctypes.cast(buf, PTR)
    null = PTR(0)

    def gather(name, *tail, e_prev=p, j=(0, 8, 8), b=(0, 1, 1), t_pad=256, ld=320, pad=64):
        fn = getattr(lib, name)
        j_lo, j_hi, n = j
        b_lo, b_hi, nb = b
        return fn(e_prev, p, p, p, *tail[:1], I64(n), I64(n), I64(1), I64(1), I64(nb),
                  I64(b_lo), I64(b_hi), I64(j_lo), I64(j_hi), I64(t_pad), I64(ld), I64(pad),
                  *tail[1:], INT(0), null)

    cases = {
        "null pointer": gather("spb_exchange_gather_tiled", null, e_prev=null),
        "receiver range": gather("spb_exchange_gather_tiled", null, j=(0, 9, 8)),
        "band range": gather("spb_exchange_gather_tiled", null, b=(1, 0, 1)),
        "multiple of the receiver tile": gather("spb_exchange_gather_tiled", null, j=(4, 8, 8)),
        "layout": gather("spb_exchange_gather_tiled", null, t_pad=250, ld=314),
        "pad": gather("spb_exchange_gather_tiled", null, ld=288, pad=32),
        "window must be": gather("spb_exchange_gather_tmem", null, I64(7)),
    }
    for text, rc in cases.items():
        assert rc < 0, text
    # the message names the violated condition
    rc = gather("spb_exchange_gather_tmem", null, I64(7))
    assert rc < 0 and b"window must be 4 or 10" in lib.spb_last_error()
    rc = gather("spb_exchange_gather_tiled", null, j=(0, 9, 8))
    assert rc < 0 and b"receiver range" in lib.spb_last_error()


def test_geometry_queries_need_no_gpu(lib_path):
    from sparrowpy_b200 import _lib, exchange
    n_r, bucket, rec = exchange.tile_geometry(_lib.F64)
    assert (n_r, bucket, rec) == (8, 32, 80)
    assert exchange.tile_geometry(_lib.F32)[2] == 48
    assert exchange.window_geometry(_lib.F64) == (8, 10, 80)
    with pytest.raises(_lib.SparrowB200Error):
        exchange.window_geometry(_lib.F32)          # the tensor-memory kernel is FP64 only
    lib = _raw(lib_path)
    t_pad, pad = I64(0), I64(0)
    assert lib.spb_exchange_layout(I64(0), I64(0), INT(0), ctypes.byref(t_pad),
                                   ctypes.byref(pad)) < 0
    assert b"n_samples" in lib.spb_last_error()
