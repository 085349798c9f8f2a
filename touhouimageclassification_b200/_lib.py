"""ctypes binding of ``csrc/libtic_b200.so`` (the C-ABI declared in ``include/tic_b200.h``).

There is no fallback: if the library is missing the import of any compute entry point raises.
"""
from __future__ import annotations

import ctypes
import os

from . import build as _build

_LIB = None

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_i64 = ctypes.c_int64
c_float = ctypes.c_float
c_u64 = ctypes.c_uint64


class TicError(RuntimeError):
    """Non-zero status from the C-ABI (message from ``tic_last_error``)."""


def lib_path() -> str:
    return _build.LIB_PATH


def load() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            raise TicError(
                f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU or PyTorch fallback for the B200 hot path)")
        _LIB = ctypes.CDLL(path)
        _LIB.tic_last_error.restype = ctypes.c_char_p
        _LIB.tic_abi_version.restype = c_int
    return _LIB


def check(status: int) -> None:
    if status != 0:
        msg = load().tic_last_error().decode("utf-8", "replace")
        raise TicError(f"tic_b200 error {status}: {msg}")


def ptr(t) -> int:
    """Device pointer of a torch tensor (or 0 for None)."""
    return 0 if t is None else t.data_ptr()


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def call(name: str, *args) -> None:
    fn = getattr(load(), name)
    check(fn(*args))
