"""Shared helpers of the energy package (``mythos/energy/utils.py:135-148`` and the model factories)."""

from __future__ import annotations

from typing import Any

import torch

from mythos_b200.input import toml


def _cast(x: Any) -> Any:
    if isinstance(x, dict):
        return {k: _cast(v) for k, v in x.items()}
    if isinstance(x, (bool, str)):
        return x
    return torch.as_tensor(x, dtype=torch.float64)


def default_configs_for(base: str) -> tuple[dict, dict]:
    """(simulation constants, energy constants) of a model as float64 tensors (``utils.py:135-148``)."""
    d = toml.load_model_defaults(base)
    sim = d.get("simulation")
    if sim is None:  # rna2 / na1 ship no simulation table in the reference either; oxDNA2's conditions apply
        sim = toml.load_model_defaults("dna2")["simulation"]
    return _cast(sim), _cast(d["energy"])
