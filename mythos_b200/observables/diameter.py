"""Helical diameter observable (``mythos/observables/diameter.py``)."""

from __future__ import annotations

import dataclasses as dc
from collections.abc import Callable

import torch

from mythos_b200 import _lib
from mythos_b200.observables import base as jd_obs

TARGETS = {"oxDNA": 23.0}  # Angstroms
ERR_DISPLACEMENT_FN_REQUIRED = "A displacement function is required for computing the helical diameter."


@dc.dataclass(frozen=True, kw_only=True)
class Diameter(jd_obs.BaseObservable):
    """Mean over the base pairs of the backbone-backbone distance plus ``sigma_backbone``, Angstrom (``diameter.py:63-76``).
    ``sigma_backbone`` may also be fixed on the object so that the observable can join an ``ObservableSet``."""

    h_bonded_base_pairs: torch.Tensor
    displacement_fn: Callable
    sigma_backbone: float | None = None

    def __post_init__(self) -> None:
        if self.rigid_body_transform_fn is None:
            raise ValueError(jd_obs.ERR_RIGID_BODY_TRANSFORM_FN_REQUIRED)
        if self.displacement_fn is None:
            raise ValueError(ERR_DISPLACEMENT_FN_REQUIRED)

    def __call__(self, trajectory, sigma_backbone: float | None = None) -> torch.Tensor:
        sigma = self.sigma_backbone if sigma_backbone is None else sigma_backbone
        if sigma is None:
            raise TypeError("Diameter needs sigma_backbone (call argument, as in the reference, or field)")
        cols = jd_obs.columns(self.rigid_body_transform_fn, self.displacement_fn, trajectory, base_pairs=self.h_bonded_base_pairs,
                              sigma_backbone=float(sigma))
        return cols[:, _lib.OBS_DIAMETER]
