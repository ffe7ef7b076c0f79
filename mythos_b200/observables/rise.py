"""Rise observable (``mythos/observables/rise.py``)."""

from __future__ import annotations

import dataclasses as dc
from collections.abc import Callable

import torch

from mythos_b200 import _lib
from mythos_b200.observables import base as jd_obs

TARGETS = {"oxDNA": 3.4}  # Angstroms


@dc.dataclass(frozen=True, kw_only=True)
class Rise(jd_obs.BaseObservable):
    """Mean over the quartets of the midpoint displacement projected on the local helical axis, Angstrom (``rise.py:58-70``)."""

    quartets: torch.Tensor
    displacement_fn: Callable

    def __post_init__(self) -> None:
        if self.rigid_body_transform_fn is None:
            raise ValueError(jd_obs.ERR_RIGID_BODY_TRANSFORM_FN_REQUIRED)

    def __call__(self, trajectory) -> torch.Tensor:
        cols = jd_obs.columns(self.rigid_body_transform_fn, self.displacement_fn, trajectory, quartets=self.quartets)
        return cols[:, _lib.OBS_RISE]
