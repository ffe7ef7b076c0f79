"""Differentiable entry to the fused CUDA energy kernels (the role ``jax.ffi.ffi_call`` + ``custom_vjp`` plays
in the XLA integration, here with torch.autograd for device memory and the chain rule).

``energy_terms(...)`` returns the per-term energies ``(F, 8)`` of ``F`` frames in one launch pair
(bonded + unbonded).  Its backward is remat-style: the forward saves only its inputs, the backward is a second
C-ABI call that recomputes the pairs and writes ``d/dcenter (F,N,3)``, ``d/dquat (F,N,4)`` and
``d/dparams (n_banks*P,)`` for the incoming cotangent ``(F, 8)`` -- exactly the differentiation contract of
SURVEY 8b.  There is no other implementation behind this function: without the shared object or without a CUDA
device it raises.
"""

from __future__ import annotations

import ctypes as C
import dataclasses as dc

import torch

from mythos_b200 import _lib


@dc.dataclass(frozen=True)
class DeviceTopology:
    """Integer inputs of the kernels, resident on the device (int32, contiguous)."""

    n: int
    seq: torch.Tensor
    bonded: torch.Tensor  # (B,2)
    nt_type: torch.Tensor | None = None
    nt_type_stack: torch.Tensor | None = None
    is_end: torch.Tensor | None = None


def _as_i32(x, device) -> torch.Tensor | None:
    if x is None:
        return None
    t = torch.as_tensor(x)
    return t.to(device=device, dtype=torch.int32).contiguous()


def make_device_topology(n, seq, bonded, device, nt_type=None, nt_type_stack=None, is_end=None) -> DeviceTopology:
    b = _as_i32(bonded, device)
    b = b.reshape(-1, 2) if b is not None else torch.zeros((0, 2), dtype=torch.int32, device=device)
    return DeviceTopology(
        n=int(n),
        seq=_as_i32(seq, device),
        bonded=b,
        nt_type=_as_i32(nt_type, device),
        nt_type_stack=_as_i32(nt_type_stack, device),
        is_end=_as_i32(is_end, device),
    )


def _launch(
    model: _lib.Model,
    topo: DeviceTopology,
    center: torch.Tensor,
    quat: torch.Tensor,
    params: torch.Tensor,
    pairs: torch.Tensor | None,
    pair_frame_stride: int,
    term_mask: int,
    cot: torch.Tensor | None,
    want_terms: bool,
    want_pos_grad: bool,
    want_param_grad: bool,
    per_frame_param_grad: bool = False,
):
    _lib.require_cuda(center, "center")
    F, N = center.shape[0], center.shape[1]
    dtype, dev = center.dtype, center.device
    sfx = _lib.suffix(dtype)
    if N != topo.n:
        raise _lib.MythosB200Error(f"body has {N} nucleotides, topology has {topo.n}")
    center = center.contiguous()
    quat = quat.to(dtype).contiguous()
    params = params.to(device=dev, dtype=dtype).contiguous()
    np_ = model.n_banks * _lib.param_count()
    if params.numel() != np_:
        raise _lib.MythosB200Error(f"params has {params.numel()} entries, expected {np_}")
    terms = torch.empty((F, _lib.N_TERMS), dtype=dtype, device=dev) if want_terms else None
    d_center = torch.empty_like(center) if want_pos_grad else None
    d_quat = torch.empty_like(quat) if want_pos_grad else None
    d_params = None
    stride = 0
    if want_param_grad:
        if per_frame_param_grad:
            d_params = torch.empty((F, np_), dtype=dtype, device=dev)
            stride = np_
        else:
            d_params = torch.empty((np_,), dtype=dtype, device=dev)
    if cot is not None:
        cot = cot.to(device=dev, dtype=dtype).contiguous()
    cap = 0
    if pairs is not None and pairs.numel() > 0:
        cap = pairs.shape[-1]
    a = _lib.EnergyArgs()
    a.model = C.pointer(model)
    a.n, a.n_frames = N, F
    a.center, a.quat = center.data_ptr(), quat.data_ptr()
    a.seq = topo.seq.data_ptr()
    a.nt_type = _lib.ptr(topo.nt_type)
    a.nt_type_stack = _lib.ptr(topo.nt_type_stack)
    a.is_end = _lib.ptr(topo.is_end)
    a.bonded = topo.bonded.data_ptr() if topo.bonded.numel() else None
    a.n_bonded = topo.bonded.shape[0]
    a.pairs = pairs.data_ptr() if cap else None
    a.pair_capacity = cap
    a.pair_frame_stride = pair_frame_stride
    a.params = params.data_ptr()
    a.cot = _lib.ptr(cot)
    a.term_mask = term_mask
    a.flags = 0
    a.terms = _lib.ptr(terms)
    a.d_center = _lib.ptr(d_center)
    a.d_quat = _lib.ptr(d_quat)
    a.d_params = _lib.ptr(d_params)
    a.d_params_frame_stride = stride
    fn = getattr(_lib.lib(), f"mythos_b200_energy_{sfx}")
    with torch.cuda.device(dev):
        _lib.check(fn(_lib.current_stream(dev), C.byref(a)), "mythos_b200_energy")
    return terms, d_center, d_quat, d_params


class _EnergyTerms(torch.autograd.Function):
    @staticmethod
    def forward(ctx, center, quat, params, model, topo, pairs, pair_frame_stride, term_mask):
        terms, _, _, _ = _launch(model, topo, center, quat, params, pairs, pair_frame_stride, term_mask, None, True, False, False)
        ctx.save_for_backward(center, quat, params)
        ctx.static = (model, topo, pairs, pair_frame_stride, term_mask)
        return terms

    @staticmethod
    def backward(ctx, g_terms):
        center, quat, params = ctx.saved_tensors
        model, topo, pairs, stride, mask = ctx.static
        need_pos = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        need_par = ctx.needs_input_grad[2]
        _, d_center, d_quat, d_params = _launch(
            model, topo, center, quat, params, pairs, stride, mask, g_terms, False, need_pos, need_par
        )
        if d_params is not None:
            d_params = d_params.to(device=params.device, dtype=params.dtype)
        return (
            d_center if ctx.needs_input_grad[0] else None,
            d_quat if ctx.needs_input_grad[1] else None,
            d_params,
            None,
            None,
            None,
            None,
            None,
        )


def energy_terms(
    model: _lib.Model,
    topo: DeviceTopology,
    center: torch.Tensor,
    quat: torch.Tensor,
    params: torch.Tensor,
    pairs: torch.Tensor | None,
    term_mask: int = _lib.ALL_TERMS,
    pair_frame_stride: int = 0,
) -> torch.Tensor:
    """Per-term energies ``(F, 8)`` for ``center (F,N,3)``, ``quat (F,N,4)``; differentiable in center, quat, params."""
    if center.dim() != 3 or quat.dim() != 3:
        raise _lib.MythosB200Error("center must be (F,N,3) and quat (F,N,4)")
    out = []
    for lo in range(0, center.shape[0], 65535):  # gridDim.y limit of one launch
        sl = slice(lo, lo + 65535)
        p = pairs
        if pairs is not None and pair_frame_stride:
            p = pairs[sl]
        out.append(_EnergyTerms.apply(center[sl], quat[sl], params, model, topo, p, pair_frame_stride, term_mask))
    return out[0] if len(out) == 1 else torch.cat(out)


def energy_and_gradients(
    model: _lib.Model,
    topo: DeviceTopology,
    center: torch.Tensor,
    quat: torch.Tensor,
    params: torch.Tensor,
    pairs: torch.Tensor | None,
    cot: torch.Tensor | None = None,
    term_mask: int = _lib.ALL_TERMS,
    pair_frame_stride: int = 0,
    want_pos_grad: bool = True,
    want_param_grad: bool = False,
    per_frame_param_grad: bool = False,
):
    """One fused launch pair returning ``(terms, d_center, d_quat, d_params)`` without autograd bookkeeping.

    This is the call the MD loop and the DiffTRe pass use: energies, forces and the parameter gradient of
    ``sum_t cot[f,t] * E_t(frame f)`` come out of a single pass over the pair list.
    """
    return _launch(
        model, topo, center, quat, params, pairs, pair_frame_stride, term_mask, cot, True, want_pos_grad, want_param_grad,
        per_frame_param_grad,
    )
