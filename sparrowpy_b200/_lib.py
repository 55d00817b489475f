"""ctypes binding of ``libsparrow_b200.so`` (the C ABI in include/sparrow_b200.h).

There is no CPU fallback: if the CUDA library is missing or no CUDA device is
present, every compute entry point raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsparrow_b200.so")
_lib = None

F64, F32 = 0, 1
_TORCH_DTYPE = {F64: torch.float64, F32: torch.float32}


class SparrowB200Error(RuntimeError):
    pass


def dtype_code(dtype):
    if dtype in (torch.float64, "f64", "float64", F64):
        return F64
    if dtype in (torch.float32, "f32", "float32"):
        return F32
    raise ValueError(f"unsupported dtype {dtype!r} (use 'f64' or 'f32')")


def torch_dtype(code):
    return _TORCH_DTYPE[code]


def load():
    """Load the CUDA library; fail loudly when it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SparrowB200Error(
                f"{LIB_PATH} not found: build it with `python -m sparrowpy_b200.build` "
                "(sparrowpy_b200 has no CPU fallback)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.spb_last_error.restype = ctypes.c_char_p
    return _lib


def require_cuda():
    if not torch.cuda.is_available():
        raise SparrowB200Error(
            "no CUDA device: sparrowpy_b200 runs on B200 GPUs only (no CPU fallback)")


def _arg(a):
    if a is None:
        return ctypes.c_void_p(0)
    if isinstance(a, torch.Tensor):
        if not a.is_cuda:
            raise SparrowB200Error("expected a CUDA tensor at the C-ABI boundary")
        if not a.is_contiguous():
            raise SparrowB200Error("expected a contiguous tensor at the C-ABI boundary")
        return ctypes.c_void_p(a.data_ptr())
    if isinstance(a, bool):
        return ctypes.c_int(int(a))
    if isinstance(a, int):
        return ctypes.c_int64(a)
    if isinstance(a, float):
        return ctypes.c_double(a)
    return a


class I32:
    """Marks a Python int that the C signature takes as `int`."""

    def __init__(self, v):
        self.v = int(v)


def _conv(a):
    if isinstance(a, I32):
        return ctypes.c_int(a.v)
    return _arg(a)


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def call(name, *args):
    """Call an ``int spb_*`` entry point; raise on a non-zero return."""
    lib = load()
    require_cuda()
    fn = getattr(lib, name)
    rc = fn(*[_conv(a) for a in args])
    if rc != 0:
        msg = lib.spb_last_error()
        raise SparrowB200Error(f"{name} failed ({rc}): {msg.decode() if msg else ''}")
    return rc


def exchange_layout(n_samples, max_delay, dtype):
    lib = load()
    t_pad, pad = ctypes.c_int64(0), ctypes.c_int64(0)
    rc = lib.spb_exchange_layout(ctypes.c_int64(n_samples), ctypes.c_int64(max_delay),
                                 ctypes.c_int(dtype), ctypes.byref(t_pad),
                                 ctypes.byref(pad))
    if rc != 0:
        raise SparrowB200Error(lib.spb_last_error().decode())
    return t_pad.value, pad.value
