"""Displacement / shift callables with the calling convention of jax_md.space (third party in the reference).

``displacement_fn(Ra, Rb) = Ra - Rb`` (free) or ``mod(Ra - Rb + L/2, L) - L/2`` (periodic); ``shift_fn(R, dR)``.
The callables carry their ``box`` so the kernels can apply the same rule on the device; calling them evaluates
the rule on torch tensors (used by observables and tests, not by the energy path).
"""

from __future__ import annotations

import torch


class Displacement:
    def __init__(self, box=None):
        self.box = None if box is None else tuple(float(b) for b in (box if hasattr(box, "__len__") else (box,) * 3))

    def box3(self) -> tuple[float, float, float]:
        return (0.0, 0.0, 0.0) if self.box is None else self.box

    def __call__(self, ra: torch.Tensor, rb: torch.Tensor) -> torch.Tensor:
        d = ra - rb
        if self.box is None:
            return d
        L = torch.as_tensor(self.box, dtype=d.dtype, device=d.device)
        return torch.remainder(d + 0.5 * L, L) - 0.5 * L

    def __eq__(self, other) -> bool:
        return isinstance(other, Displacement) and self.box == other.box

    def __hash__(self) -> int:
        return hash(("Displacement", self.box))

    def __repr__(self) -> str:
        return f"Displacement(box={self.box})"


class Shift:
    def __init__(self, box=None):
        self.box = Displacement(box).box

    def __call__(self, r: torch.Tensor, dr: torch.Tensor) -> torch.Tensor:
        if self.box is None:
            return r + dr
        L = torch.as_tensor(self.box, dtype=r.dtype, device=r.device)
        return torch.remainder(r + dr, L)


def free() -> tuple[Displacement, Shift]:
    return Displacement(None), Shift(None)


def periodic(box) -> tuple[Displacement, Shift]:
    return Displacement(box), Shift(box)


def box_of(displacement_fn) -> tuple[float, float, float]:
    """Box of a displacement callable made by this module (free space -> zeros)."""
    if isinstance(displacement_fn, Displacement):
        return displacement_fn.box3()
    box = getattr(displacement_fn, "box", None)
    if box is None:
        raise TypeError(
            "displacement_fn must come from mythos_b200.space.free()/periodic(): the kernels need the box, "
            "an opaque callable cannot be traced into CUDA"
        )
    return Displacement(box).box3()
