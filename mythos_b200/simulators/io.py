"""Trajectory container (behaviour of ``mythos/simulators/io.py:19-225``).

``SimulatorTrajectory`` is a stacked RigidBody -- ``center (F,N,3)``, ``orientation.vec (F,N,4)`` -- plus optional
per-state ``temperature`` (kT), ``box_size`` and ``metadata``.  ``length / slice / filter / with_state_metadata /
concat / __add__ / from_rigid_body`` behave as in the reference: ``slice`` slices every per-state field (box sizes
and metadata included), ``concat`` raises on an empty list, refuses a mix of present and absent optional fields and
merges metadata key by key (missing keys are filled with NaN rows).  Tensors may live on the host (pinned, for ingest)
or on the device.

``shard = (lo, hi, total)`` marks a block of a frame-sharded trajectory (multi-GPU DiffTRe): such a block cannot be
sliced or concatenated without invalidating the bounds, so both operations return an un-sharded result only when they
keep the block whole and raise otherwise.
"""

from __future__ import annotations

import dataclasses as dc
from collections.abc import Callable
from typing import Any

import torch

from mythos_b200.rigid_body import Quaternion, RigidBody


def _concat_optional_field(values: list, label: str):
    """None if every entry is None, error on a mix, concatenation otherwise (``io.py:175-188``)."""
    if all(v is None for v in values):
        return None
    if any(v is None for v in values):
        raise ValueError(f"Cannot concatenate, trajectories have incompatible {label}.")
    return torch.cat([torch.as_tensor(v) for v in values], dim=0)


def _merge_metadata(metadata_list: list, lengths: list[int]):
    """Key-wise merge for ``concat`` (``io.py:191-217``): a key missing from some trajectories is filled with NaN rows
    of the shape the others have; shapes beyond the leading axis must agree."""
    if all(not m for m in metadata_list):
        return None
    dicts = [dict(m) if m else {} for m in metadata_list]
    for key in {k for d in dicts for k in d}:
        present = [torch.as_tensor(d[key]) for d in dicts if key in d]
        shape = tuple(present[0].shape[1:])
        if any(tuple(p.shape[1:]) != shape for p in present[1:]):
            raise ValueError(f"Metadata key '{key}' has mismatched shapes when adding trajectories.")
        proto = present[0]
        dtype = proto.dtype if proto.dtype.is_floating_point else torch.float64
        for d, length in zip(dicts, lengths, strict=True):
            if key not in d:
                d[key] = torch.full((length, *shape), float("nan"), dtype=dtype, device=proto.device)
    return {k: torch.cat([torch.as_tensor(d[k]) for d in dicts], dim=0) for k in dicts[0]}


@dc.dataclass(frozen=True)
class SimulatorTrajectory(RigidBody):
    box_size: Any = None
    temperature: torch.Tensor | None = None
    metadata: dict[str, torch.Tensor] | None = None
    shard: tuple[int, int, int] | None = None  # (lo, hi, total): these states are frames lo..hi of a sharded trajectory

    @classmethod
    def from_rigid_body(cls, rigid_body: RigidBody, **kwargs: Any) -> "SimulatorTrajectory":
        return cls(center=rigid_body.center, orientation=rigid_body.orientation, **kwargs)

    def replace(self, **changes: Any) -> "SimulatorTrajectory":
        return dc.replace(self, **changes)

    def with_state_metadata(self, **metadata: Any) -> "SimulatorTrajectory":
        """Set the same metadata for all states in the trajectory (``io.py:62-67``)."""
        new = dict(self.metadata) if self.metadata is not None else {}
        for key, value in metadata.items():
            new[key] = torch.stack([torch.as_tensor(value)] * self.length())
        return dc.replace(self, metadata=new)

    def filter(self, filter_fn: Callable[[Any], Any]) -> "SimulatorTrajectory":
        """Keep the states for which ``filter_fn(metadata)`` is true (``io.py:69-81``)."""
        keep = torch.as_tensor(filter_fn(self.metadata))
        return self.slice(torch.where(keep)[0])

    def length(self) -> int:
        return int(self.center.shape[0])

    def slice(self, key) -> "SimulatorTrajectory":
        """Slice every per-state field: centres, orientations, box sizes, temperatures, metadata (``io.py:83-103``)."""
        if isinstance(key, int):
            key = slice(key, key + 1)
        if not isinstance(key, slice):
            key = torch.as_tensor(key)
        center = self.center[key]
        shard = self.shard
        if shard is not None and center.shape[0] != self.center.shape[0]:
            raise ValueError(
                "cannot slice a frame-sharded trajectory block (shard bounds would no longer describe it): slice the "
                "full trajectory before sharding"
            )

        def per_state(x):
            return None if x is None else (x[key.to(x.device)] if isinstance(key, torch.Tensor) and isinstance(x, torch.Tensor) else x[key])

        box = self.box_size
        if box is not None and hasattr(box, "__getitem__") and getattr(box, "ndim", 0) >= 1 and len(box) == self.length():
            box = per_state(box)
        return dc.replace(
            self,
            center=center,
            orientation=Quaternion(per_state(self.orientation.vec)),
            box_size=box,
            temperature=per_state(self.temperature),
            metadata=None if self.metadata is None else {k: per_state(v) for k, v in self.metadata.items()},
            shard=shard,
        )

    def to(self, *args, **kwargs) -> "SimulatorTrajectory":
        def mv(x):
            return x.to(*args, **kwargs) if isinstance(x, torch.Tensor) else x

        return dc.replace(
            self,
            center=mv(self.center),
            orientation=Quaternion(mv(self.orientation.vec)),
            box_size=mv(self.box_size),
            temperature=mv(self.temperature),
            metadata=None if self.metadata is None else {k: mv(v) for k, v in self.metadata.items()},
        )

    @classmethod
    def concat(cls, trajectories: list["SimulatorTrajectory"]) -> "SimulatorTrajectory":
        """Concatenate along the state axis (``io.py:117-143``)."""
        if not trajectories:
            raise ValueError("Cannot concatenate an empty list of trajectories.")
        if len(trajectories) == 1:
            return trajectories[0]
        if any(t.shard is not None for t in trajectories):
            raise ValueError("cannot concatenate frame-sharded trajectory blocks: concatenate before sharding")
        box_size = _concat_optional_field([t.box_size for t in trajectories], "box sizes")
        temperature = _concat_optional_field([t.temperature for t in trajectories], "temperatures")
        metadata = _merge_metadata([t.metadata for t in trajectories], [t.length() for t in trajectories])
        return dc.replace(
            trajectories[0],
            center=torch.cat([t.center for t in trajectories], dim=0),
            orientation=Quaternion(torch.cat([t.orientation.vec for t in trajectories], dim=0)),
            box_size=box_size,
            temperature=temperature,
            metadata=metadata,
        )

    def __add__(self, other: "SimulatorTrajectory") -> "SimulatorTrajectory":
        return self.__class__.concat([self, other])
