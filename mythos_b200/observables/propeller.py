"""Propeller twist observable (``mythos/observables/propeller.py``)."""

from __future__ import annotations

import dataclasses as dc

import torch

from mythos_b200 import _lib
from mythos_b200.observables import base as jd_obs

TARGETS = {"oxDNA": 21.7}  # degrees


@dc.dataclass(frozen=True)
class PropellerTwist(jd_obs.BaseObservable):
    """Mean over the h-bonded base pairs of ``180 - acos(clamp(n_i . n_j))`` in degrees, per state (``propeller.py:57-71``)."""

    h_bonded_base_pairs: torch.Tensor = None

    def __post_init__(self) -> None:
        if self.rigid_body_transform_fn is None:
            raise ValueError(jd_obs.ERR_RIGID_BODY_TRANSFORM_FN_REQUIRED)

    def __call__(self, trajectory) -> torch.Tensor:
        cols = jd_obs.columns(self.rigid_body_transform_fn, None, trajectory, base_pairs=self.h_bonded_base_pairs)
        return cols[:, _lib.OBS_PROPELLER]
