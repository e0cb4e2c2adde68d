"""ctypes binding of the C-ABI library ``libimagine360_b200.so`` (declared in include/imagine360_b200.h).

The product path has no CPU fallback: if the library is missing, importing any op fails loudly.
"""
from __future__ import annotations

import ctypes
import os
import sys
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ["I360_LIB_PATH"]) if os.environ.get("I360_LIB_PATH") else _HERE / "libimagine360_b200.so"

_lib = None

ERRORS = {
    -1: "I360_ERR_ARG (bad shape / alignment / null pointer)",
    -2: "I360_ERR_CUDA (CUDA runtime error)",
    -3: "I360_ERR_TMAP (TMA descriptor creation failed)",
    -4: "I360_ERR_UNSUPPORTED",
}


class NativeLibraryMissing(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise NativeLibraryMissing(
                f"{LIB_PATH} not found: build it with `python -m imagine360_b200.build` "
                "(there is no CPU fallback for the product path)"
            )
        _lib = ctypes.CDLL(str(LIB_PATH))
    return _lib


LAUNCHES = 0   # C-ABI kernel launches issued through this module (bench.py reports it as gpu_launches)


def check(code: int, what: str) -> None:
    """Translate a C-ABI status into an exception.

    The reference wraps the pipeline call in a bare ``except: continue``
    (inference_dual_p2e.py:596-597), so also write to stderr before raising.
    """
    global LAUNCHES
    LAUNCHES += 1
    if code != 0:
        msg = f"imagine360_b200: {what} failed: {ERRORS.get(code, code)}"
        print(msg, file=sys.stderr, flush=True)
        raise RuntimeError(msg)


def exported_symbols() -> list[str]:
    """Names declared in include/imagine360_b200.h (parsed from the header)."""
    import re

    hdr = (_HERE.parent / "include" / "imagine360_b200.h").read_text()
    return sorted(set(re.findall(r"\b(i360_[a-z0-9_]+)\s*\(", hdr)))
