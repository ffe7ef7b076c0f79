"""Minimal rigid-body containers with the field names the reference takes from jax_md.rigid_body.

``RigidBody(center, orientation=Quaternion(vec))`` with ``center`` (N,3) or (F,N,3) and ``vec`` (N,4) or
(F,N,4) = (w,x,y,z) torch tensors.  They are plain holders: device memory and autograd
plumbing only; no arithmetic lives here.
"""

from __future__ import annotations

import dataclasses as dc

import torch


@dc.dataclass(frozen=True)
class Quaternion:
    vec: torch.Tensor

    def __getitem__(self, key) -> "Quaternion":
        return Quaternion(self.vec[key])


@dc.dataclass(frozen=True)
class RigidBody:
    center: torch.Tensor
    orientation: Quaternion

    def __getitem__(self, key) -> "RigidBody":
        return RigidBody(self.center[key], self.orientation[key])

    @property
    def n_frames(self) -> int | None:
        return self.center.shape[0] if self.center.dim() == 3 else None

    def to(self, *args, **kwargs) -> "RigidBody":
        return RigidBody(self.center.to(*args, **kwargs), Quaternion(self.orientation.vec.to(*args, **kwargs)))

    def detach(self) -> "RigidBody":
        return RigidBody(self.center.detach(), Quaternion(self.orientation.vec.detach()))

    def requires_grad_(self, flag: bool = True) -> "RigidBody":
        self.center.requires_grad_(flag)
        self.orientation.vec.requires_grad_(flag)
        return self


def stack(bodies: list[RigidBody]) -> RigidBody:
    return RigidBody(torch.stack([b.center for b in bodies]), Quaternion(torch.stack([b.orientation.vec for b in bodies])))
