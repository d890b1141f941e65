"""ctypes binding of libwhisper_b200.so - the only route from Python to the CUDA kernels.

There is deliberately no fallback: if the shared library is missing or a symbol cannot be bound,
importing this module's `lib()` raises, and every operator in the package fails with it.
"""
from __future__ import annotations

import ctypes
import os
import re
from ctypes import c_char_p, c_float, c_int, c_int64, c_uint64, c_void_p
from typing import Dict, List, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libwhisper_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "whisper_b200.h")

_lib = None


class WhisperB200Error(RuntimeError):
    pass


EXPORTS_PATH = os.path.join(_HERE, "lib", "exports.txt")


def header_symbols() -> List[str]:
    """Every function name declared in include/whisper_b200.h (used by the ABI export test).  A copy of the package
    without the repository's include/ directory falls back to lib/exports.txt, the list build.py writes next to the
    shared library from that same header."""
    if os.path.exists(HEADER_PATH):
        text = open(HEADER_PATH).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        return sorted(set(re.findall(r"\b(wb200_[a-z0-9_]+)\s*\(", text)))
    with open(EXPORTS_PATH) as f:
        return sorted(set(f.read().split()))


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise WhisperB200Error(
                f"{LIB_PATH} not found: build it with `python -m whisper_b200.build` "
                "(there is no CPU or PyTorch fallback for the hot path)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.wb200_version.restype = c_char_p
        _lib.wb200_last_error.restype = c_char_p
        _lib.wb200_launch_count.restype = c_uint64
        for name in header_symbols():
            fn = getattr(_lib, name)  # raises AttributeError if the export is missing
            if name in ("wb200_version", "wb200_last_error", "wb200_launch_count"):
                continue
            if name.endswith("_bytes"):
                fn.restype = ctypes.c_size_t
            elif name == "wb200_decoder_logits_ld":
                fn.restype = c_int64
            elif name.endswith("_destroy"):
                fn.restype = None
            else:
                fn.restype = c_int
    return _lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = lib().wb200_last_error().decode()
        raise WhisperB200Error(f"{what or 'whisper_b200'} failed with status {status}: {msg}")


def launch_count() -> int:
    return int(lib().wb200_launch_count())


def ptr(t) -> c_void_p:
    """Device (or host) pointer of a torch tensor / None."""
    if t is None:
        return c_void_p(0)
    return c_void_p(t.data_ptr())


def stream_ptr() -> c_void_p:
    import torch

    return c_void_p(torch.cuda.current_stream().cuda_stream)


def dtype_code(dtype) -> int:
    import torch

    if dtype == torch.bfloat16:
        return 0
    if dtype == torch.float16:
        return 1
    raise WhisperB200Error(f"unsupported activation dtype {dtype}; use torch.bfloat16 or torch.float16")
