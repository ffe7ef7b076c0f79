"""Test helper: compiles mythos_b200/csrc/oxdna_device.cuh for the HOST (g++) and evaluates it pair by pair.

TEST INFRASTRUCTURE ONLY -- this is how the analytic gradients are checked in the CPU suite (no GPU here);
the shipped library has no host path and nothing under mythos_b200/ imports this module.
"""

from __future__ import annotations

import ctypes as C
import subprocess
from functools import lru_cache
from pathlib import Path

import numpy as np

from mythos_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "tests" / "host_check" / "host_check.cpp"
OUT = ROOT / "tests" / "_build" / "libhost_check.so"


@lru_cache(maxsize=1)
def lib() -> C.CDLL:
    deps = [SRC, ROOT / "mythos_b200" / "csrc" / "oxdna_device.cuh", ROOT / "include" / "mythos_b200.h"]
    if not OUT.exists() or any(d.stat().st_mtime > OUT.stat().st_mtime for d in deps):
        OUT.parent.mkdir(parents=True, exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", str(OUT), str(SRC)], check=True)
    h = C.CDLL(str(OUT))
    h.host_check_eval.restype = None
    return h


def _p(a, ctype):
    return None if a is None else a.ctypes.data_as(C.POINTER(ctype))


def evaluate(model: _lib.Model, center, quat, seq, bonded, pairs, params, cot=None, mask=0xFF, nt_type=None,
             nt_type_stack=None, is_end=None, use_f32=False):
    """-> terms (8,), d_center (N,3), d_quat (N,4), d_params (n_banks*P,) of sum_t cot_t E_t for ONE frame."""
    n = center.shape[0]
    f64 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)  # noqa: E731
    i32 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.int32)  # noqa: E731
    center, quat, params, cot = f64(center), f64(quat), f64(params), f64(cot)
    seq, nt_type, nt_type_stack, is_end = i32(seq), i32(nt_type), i32(nt_type_stack), i32(is_end)
    bonded = i32(np.asarray(bonded).reshape(-1, 2))
    pairs = i32(np.asarray(pairs).reshape(2, -1))
    terms = np.zeros(8)
    dc_, dq = np.zeros((n, 3)), np.zeros((n, 4))
    dp = np.zeros(model.n_banks * _lib.param_count())
    lib().host_check_eval(
        C.c_int(1 if use_f32 else 0), C.byref(model), C.c_int(n), _p(center, C.c_double), _p(quat, C.c_double),
        _p(seq, C.c_int), _p(nt_type, C.c_int), _p(nt_type_stack, C.c_int), _p(is_end, C.c_int),
        _p(bonded, C.c_int), C.c_int(bonded.shape[0]), _p(pairs, C.c_int), C.c_long(pairs.shape[1]),
        _p(params, C.c_double), _p(cot, C.c_double), C.c_uint(mask), _p(terms, C.c_double), _p(dc_, C.c_double),
        _p(dq, C.c_double), _p(dp, C.c_double),
    )
    return terms, dc_, dq, dp


def scalar(kind: int, x: float, block, eps: float = 1.0, use_f32: bool = False) -> tuple[float, float]:
    """(value, d/dx) of the device header's f<kind>_val at x for a parameter block in the header's layout."""
    h = lib()
    h.host_check_scalar.restype = C.c_double
    p = np.zeros(16)
    p[: len(block)] = block
    df = C.c_double(0.0)
    v = h.host_check_scalar(C.c_int(1 if use_f32 else 0), C.c_int(kind), C.c_double(x), _p(p, C.c_double), C.c_double(eps), C.byref(df))
    return float(v), float(df.value)
