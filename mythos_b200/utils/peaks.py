"""Roofline denominators measured on the box: FMA issue peak and the issue cost of the special functions the kernels call
(library micro-benchmarks in csrc/peaks.cu, timed with CUDA events on the launching stream)."""

from __future__ import annotations

import torch

from mythos_b200 import _lib

SPECIAL_KINDS = {"div": 0, "sqrt": 1, "exp": 2, "log": 3, "acos": 4, "rsqrt": 5, "rcp": 6, "cos": 7, "fmod": 8}
_BLOCKS = 148 * 32


def _time(fn, reps: int) -> float:
    best = float("inf")
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e-3


def fma_peak_tflops(dev, dtype=torch.float64, iters: int = 4096, reps: int = 4) -> float:
    """Measured FMA issue peak in TFLOP/s (2 flop per FMA): 148 x 32 blocks x 256 threads x 8 independent chains."""
    scratch = torch.empty(_BLOCKS * 256, dtype=dtype, device=dev)
    fn = getattr(_lib.lib(), f"mythos_b200_fma_peak_{_lib.suffix(dtype)}")
    with torch.cuda.device(dev):
        s = _time(lambda: _lib.check(fn(_lib.current_stream(dev), scratch.data_ptr(), _BLOCKS, iters), "fma_peak"), reps)
    return _BLOCKS * 256 * iters * 16 / s / 1e12


def special_weights(dev, dtype=torch.float64, iters: int = 512, reps: int = 3) -> dict:
    """Issue cost of each special function in FMA slots: t(chain of f(x)*a+b) / t(chain of FMA) - 1, both measured now.
    -> {"fma_per_s": ..., "weights": {name: slots}, "ops_per_s": {name: rate}}"""
    scratch = torch.empty(_BLOCKS * 256, dtype=dtype, device=dev)
    sfx = _lib.suffix(dtype)
    lib = _lib.lib()
    fma = getattr(lib, f"mythos_b200_fma_peak_{sfx}")
    spc = getattr(lib, f"mythos_b200_special_rate_{sfx}")
    n_ops = _BLOCKS * 256 * iters * 8
    with torch.cuda.device(dev):
        st = _lib.current_stream(dev)
        t_fma = _time(lambda: _lib.check(fma(st, scratch.data_ptr(), _BLOCKS, iters), "fma_peak"), reps + 1)
        out = {"fma_per_s": n_ops / t_fma, "weights": {}, "ops_per_s": {}}
        for name, kind in SPECIAL_KINDS.items():
            t = _time(lambda: _lib.check(spc(st, scratch.data_ptr(), _BLOCKS, iters, kind), "special_rate"), reps)
            out["weights"][name] = t / t_fma - 1.0
            out["ops_per_s"][name] = n_ops / t
    return out
