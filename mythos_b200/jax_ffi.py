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
XLA_TARGETS = tuple(f"mythos_b200_xla_{op}_{sfx}" for op in ("energy", "nl_build", "langevin", "weights_neff") for sfx in ("f64", "f32"))


def register() -> None:  # pragma: no cover
    """``jax.ffi.register_ffi_target`` for both dtypes (platform CUDA)."""
    if not HAVE_JAX:
        raise RuntimeError("jax is not installed; use mythos_b200.energy (torch) which binds the same C-ABI")
    lib = ctypes.CDLL(str(XLA_LIB))
    for name in XLA_TARGETS:
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


def _sfx(dtype) -> str:  # pragma: no cover
    return "f64" if dtype == jnp.float64 else "f32"


def neighbor_list(center, bonded, box, r_cutoff: float, dr_threshold: float, capacity: int):  # pragma: no cover
    """``jax_md.partition.neighbor_list(..., format=OrderedSparse)`` as ``mythos/utils/neighbors.py:51-59`` calls it:
    ``center (F,N,3)`` -> ``(idx (F,2,capacity) int32 padded with N, count (F), did_buffer_overflow (1))``."""
    register()
    import numpy as np

    from mythos_b200 import _lib

    F, N = center.shape[0], center.shape[1]
    ws = int(_lib.lib().mythos_b200_nl_workspace_bytes(N, F))
    out = (jax.ShapeDtypeStruct((F, 2, capacity), jnp.int32), jax.ShapeDtypeStruct((F,), jnp.int32),
           jax.ShapeDtypeStruct((1,), jnp.int32), jax.ShapeDtypeStruct((ws,), jnp.uint8))
    idx, count, overflow, _ = jax.ffi.ffi_call(f"mythos_b200_xla_nl_build_{_sfx(center.dtype)}", out, vmap_method="sequential")(
        center, bonded, box=np.asarray(box, dtype=np.float64), r_cutoff=np.float64(r_cutoff), dr_threshold=np.float64(dr_threshold))
    return idx, count, overflow


def langevin_step(state, forces, dt, kT, gamma, mass, inertia, box, seed: int, step: int, phase: int = 0, noise=None):  # pragma: no cover
    """One fused B-A-O-A (+B) sub-step on ``state = (center, quat, p_center, p_quat)`` with ``forces = (dE/dcenter, dE/dquat)``;
    the state results alias the operands (``input_output_aliases``), so under ``jit`` with donated state nothing is copied."""
    register()
    import numpy as np

    center, quat, p_center, p_quat = state
    noise = jnp.zeros((0,), center.dtype) if noise is None else noise
    out = tuple(jax.ShapeDtypeStruct(x.shape, x.dtype) for x in state)
    return jax.ffi.ffi_call(f"mythos_b200_xla_langevin_{_sfx(center.dtype)}", out, input_output_aliases={0: 0, 1: 1, 2: 2, 3: 3})(
        center, quat, p_center, p_quat, forces[0], forces[1], noise, dt=np.float64(dt), kT=np.float64(kT),
        gamma_center=np.float64(gamma[0]), gamma_quat=np.float64(gamma[1]), mass=np.float64(mass),
        inertia=np.asarray(inertia, dtype=np.float64), box=np.asarray(box, dtype=np.float64), seed=np.int64(seed), step=np.int64(step),
        phase=np.int32(phase))


def weights_and_neff(beta, new_energies, ref_energies):  # pragma: no cover
    """``compute_weights_and_neff`` (``mythos/optimization/objective.py:139-163``) with its analytic VJP."""
    register()

    @jax.custom_vjp
    def f(beta, e_new, e_ref):
        out = (jax.ShapeDtypeStruct(e_new.shape, e_new.dtype), jax.ShapeDtypeStruct((4,), e_new.dtype))
        w, sums = jax.ffi.ffi_call(f"mythos_b200_xla_weights_neff_{_sfx(e_new.dtype)}", out)(beta, e_new, e_ref)
        return w, sums[3]

    def fwd(beta, e_new, e_ref):
        w, neff = f(beta, e_new, e_ref)
        return (w, neff), (w, neff, beta)

    def bwd(res, g):
        w, neff, beta = res
        g_w, g_neff = g
        gx = w * (g_w - jnp.sum(g_w * w))
        wlw = jnp.where(w > 0, w * jnp.log(jnp.where(w > 0, w, 1.0)), 0.0)
        gx = gx + g_neff * neff * (-(wlw - w * jnp.sum(wlw)))
        g_e = -beta * gx
        return jnp.zeros_like(beta), g_e, -g_e

    f.defvjp(fwd, bwd)
    return f(beta, new_energies, ref_energies)
