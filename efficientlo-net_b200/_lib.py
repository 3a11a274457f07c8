"""ctypes binding of libelo_b200.so (the C ABI declared in include/elo_b200.h).

There is no fallback: if the CUDA library is missing or a call fails, an exception is raised.
"""
import ctypes
import os

from .build import LIB

_c_int, _c_float, _c_void_p = ctypes.c_int, ctypes.c_float, ctypes.c_void_p

# name -> argtypes; every function returns int (0 = ok), see include/elo_b200.h
_FUSED_CONV = [_c_int] * 8 + [_c_float, _c_int, _c_int] + [_c_void_p] * 8 + [_c_int, _c_int, _c_void_p]
SIGNATURES = {
    "elo_fused_conv_select_k": _FUSED_CONV,
    "elo_fused_conv_random_k": _FUSED_CONV,
}

_lib = None


class EloError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            raise EloError(
                "%s is missing: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
                "There is no CPU or PyTorch fallback for this path." % LIB)
        handle = ctypes.CDLL(LIB)
        handle.elo_last_error.restype = ctypes.c_char_p
        handle.elo_last_error.argtypes = []
        handle.elo_version.restype = _c_int
        for name, argtypes in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.argtypes = argtypes
            fn.restype = _c_int
        _lib = handle
    return _lib


def check(rc, what):
    """Map the C ABI's status to the exceptions the reference op raises (InvalidArgument -> ValueError)."""
    if rc == 0:
        return
    msg = lib().elo_last_error().decode("utf-8", "replace")
    if rc == -1:
        raise ValueError("%s: %s" % (what, msg))
    raise EloError("%s failed (status %d): %s" % (what, rc, msg))


def stream_ptr(device):
    import torch
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t):
    return None if t is None else t.data_ptr()
