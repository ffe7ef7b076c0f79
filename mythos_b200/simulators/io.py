"""Trajectory container (layout contract of ``mythos/simulators/io.py:19-170``).

``SimulatorTrajectory`` is a stacked RigidBody -- ``center (F,N,3)``, ``orientation.vec (F,N,4)`` -- plus optional
per-state ``temperature`` (kT), ``box_size`` and ``metadata``; ``length / slice / concat / from_rigid_body`` behave as
in the reference.  Tensors may live on the host (pinned, for ingest) or on the device.
"""

from __future__ import annotations

import dataclasses as dc
from typing import Any

import torch

from mythos_b200.rigid_body import Quaternion, RigidBody


@dc.dataclass(frozen=True)
class SimulatorTrajectory(RigidBody):
    box_size: Any = None
    temperature: torch.Tensor | None = None
    metadata: dict[str, torch.Tensor] | None = None
    shard: tuple[int, int, int] | None = None  # (lo, hi, total): these states are frames lo..hi of a sharded trajectory

    @classmethod
    def from_rigid_body(cls, rigid_body: RigidBody, **kwargs: Any) -> "SimulatorTrajectory":
        return cls(center=rigid_body.center, orientation=rigid_body.orientation, **kwargs)

    def length(self) -> int:
        return int(self.center.shape[0])

    def slice(self, key: int | slice) -> "SimulatorTrajectory":
        if isinstance(key, int):
            key = slice(key, key + 1)
        return dc.replace(
            self,
            center=self.center[key],
            orientation=Quaternion(self.orientation.vec[key]),
            temperature=None if self.temperature is None else self.temperature[key],
            metadata=None if self.metadata is None else {k: v[key] for k, v in self.metadata.items()},
        )

    def to(self, *args, **kwargs) -> "SimulatorTrajectory":
        return dc.replace(
            self,
            center=self.center.to(*args, **kwargs),
            orientation=Quaternion(self.orientation.vec.to(*args, **kwargs)),
            temperature=None if self.temperature is None else self.temperature.to(*args, **kwargs),
        )

    @classmethod
    def concat(cls, trajectories: list["SimulatorTrajectory"]) -> "SimulatorTrajectory":
        if len(trajectories) == 1:
            return trajectories[0]
        temps = [t.temperature for t in trajectories]
        temperature = None if any(t is None for t in temps) else torch.cat(temps)
        return cls(
            center=torch.cat([t.center for t in trajectories]),
            orientation=Quaternion(torch.cat([t.orientation.vec for t in trajectories])),
            box_size=trajectories[0].box_size,
            temperature=temperature,
        )
