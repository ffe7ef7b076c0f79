"""Pitch observable (``mythos/observables/pitch.py``)."""

from __future__ import annotations

import dataclasses as dc
import math
from collections.abc import Callable

import torch

from mythos_b200 import _lib
from mythos_b200.observables import base as jd_obs

TARGETS = {"oxDNA": 10.5}  # bp/turn


def compute_pitch(avg_pitch_angle):
    """Pitch in base pairs per turn from the trajectory-averaged pitch angle in radians (``pitch.py:19-29``)."""
    return math.pi / avg_pitch_angle


@dc.dataclass(frozen=True, kw_only=True)
class PitchAngle(jd_obs.BaseObservable):
    """Mean over the quartets of the angle between adjacent base pairs' backbone vectors, radians (``pitch.py:76-88``)."""

    quartets: torch.Tensor
    displacement_fn: Callable

    def __post_init__(self) -> None:
        if self.rigid_body_transform_fn is None:
            raise ValueError(jd_obs.ERR_RIGID_BODY_TRANSFORM_FN_REQUIRED)

    def __call__(self, trajectory) -> torch.Tensor:
        cols = jd_obs.columns(self.rigid_body_transform_fn, self.displacement_fn, trajectory, quartets=self.quartets)
        return cols[:, _lib.OBS_PITCH_ANGLE]
