"""Build/load helper for the C part of the oracle -- TEST INFRASTRUCTURE ONLY."""

from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "csrc", "sir_oracle.c")
_OUT_DIR = os.path.join(_HERE, "_build")
_OUT = os.path.join(_OUT_DIR, "libsir_oracle.so")


def build_c_oracle(force: bool = False) -> str:
    """Compile ``oracle/csrc/sir_oracle.c`` with gcc into ``oracle/_build/``."""
    os.makedirs(_OUT_DIR, exist_ok=True)
    if force or not os.path.exists(_OUT) or os.path.getmtime(_OUT) < os.path.getmtime(_SRC):
        subprocess.check_call(
            ["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", _OUT, _SRC])
    return _OUT


_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_c_oracle())
        _lib.oracle_systematic_ancestors.restype = ctypes.c_int
        _lib.oracle_systematic_ancestors.argtypes = [
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_double, ctypes.c_int, ctypes.c_void_p]
        _lib.oracle_cumsum.restype = None
        _lib.oracle_cumsum.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
    return _lib


def c_systematic_ancestors(weights: np.ndarray, u: float, clamp_last: bool = True) -> np.ndarray:
    """Literal two-pointer resampler in C (reference particle_filter.py:146-171)."""
    w = np.ascontiguousarray(weights, dtype=np.float64)
    out = np.empty(w.shape[0], dtype=np.int64)
    rc = _load().oracle_systematic_ancestors(
        w.ctypes.data, w.shape[0], float(u), int(bool(clamp_last)), out.ctypes.data)
    if rc != 0:
        raise MemoryError("oracle_systematic_ancestors: allocation failed")
    return out
