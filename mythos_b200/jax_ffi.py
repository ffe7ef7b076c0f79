"""Registration of the kernels as XLA custom calls (``jax.ffi``) with a ``jax.custom_vjp`` -- the integration the
north star names, for environments where JAX exists.

This build image has no jax / jaxlib, so this module is import-guarded and has NOT been executed
in this repository's test runs; parity and performance evidence is carried by the C-ABI through ctypes
(``mythos_b200._lib``) and torch.autograd (``mythos_b200.energy.functional``), which bind exactly the same symbols.

Usage where JAX is installed (after building ``libmythos_b200_xla.so`` as described in ``csrc/xla_ffi_shim.cc``)::

    from mythos_b200 import jax_ffi
    energy = jax_ffi.make_energy_fn(model_bytes, seq, nt_type, is_end, bonded, term_mask=0xFF)
    terms = energy(center, quat, params, pairs)                        # (F, 8), jit / grad / vmap-able
    g = jax.grad(lambda p: energy(center, quat, p, pairs).sum())(params)
"""

from __future__ import annotations

import ctypes
from pathlib import Path

try:  # pragma: no cover - jax is absent from the build image
    import jax
    import jax.numpy as jnp

    HAVE_JAX = True
except ImportError:  # pragma: no cover
    HAVE_JAX = False

XLA_LIB = Path(__file__).resolve().parent / "libmythos_b200_xla.so"


def register() -> None:  # pragma: no cover
    """``jax.ffi.register_ffi_target`` for both dtypes (platform CUDA)."""
    if not HAVE_JAX:
        raise RuntimeError("jax is not installed; use mythos_b200.energy (torch) which binds the same C-ABI")
    lib = ctypes.CDLL(str(XLA_LIB))
    for name in ("mythos_b200_xla_energy_f64", "mythos_b200_xla_energy_f32"):
        jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(getattr(lib, name)), platform="CUDA")


def make_energy_fn(model_bytes: bytes, seq, nt_type, is_end, bonded, term_mask: int = 0xFF):  # pragma: no cover
    """Per-term energies ``(F, 8)`` as a differentiable JAX function of (center, quat, params); pairs are static data."""
    if not HAVE_JAX:
        raise RuntimeError("jax is not installed")
    register()
    import numpy as np

    model = np.frombuffer(model_bytes, dtype=np.uint8)

    def call(center, quat, params, pairs, cot, want_grads):
        F, N = center.shape[0], center.shape[1]
        sfx = "f64" if center.dtype == jnp.float64 else "f32"
        out = (
            jax.ShapeDtypeStruct((F, 8), center.dtype),
            jax.ShapeDtypeStruct(center.shape if want_grads else (0,), center.dtype),
            jax.ShapeDtypeStruct(quat.shape if want_grads else (0,), center.dtype),
            jax.ShapeDtypeStruct(params.shape if want_grads else (0,), center.dtype),
        )
        return jax.ffi.ffi_call(f"mythos_b200_xla_energy_{sfx}", out, vmap_method="sequential")(
            center, quat, params, cot, seq, nt_type, is_end, bonded, pairs, model=model, term_mask=np.int32(term_mask),
            want_grads=np.int32(1 if want_grads else 0))

    @jax.custom_vjp
    def energy(center, quat, params, pairs):
        return call(center, quat, params, pairs, jnp.zeros((0,), center.dtype), False)[0]

    def fwd(center, quat, params, pairs):  # remat-safe: only inputs are saved
        return energy(center, quat, params, pairs), (center, quat, params, pairs)

    def bwd(res, g):
        center, quat, params, pairs = res
        _, dc, dq, dp = call(center, quat, params, pairs, g, True)
        return dc, dq, dp, None

    energy.defvjp(fwd, bwd)
    return energy
