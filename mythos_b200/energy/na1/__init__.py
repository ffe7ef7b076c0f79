"""oxNA (DNA/RNA hybrid) energy model (interface of ``mythos.energy.na1``).

Every configuration carries ``nt_type`` plus ``dna_*``, ``rna_*`` and (unbonded terms) ``drh_*`` copies of the
single-model parameters; ``init_params`` builds the per-bank sub-configurations ``dna_config`` / ``rna_config`` /
``drh_config`` exactly as ``mythos/energy/na1/*.py`` do.  The reference then evaluates all four variants for every
pair and ``where``-selects; the kernels select one parameter bank and one site geometry per pair
from ``nt_type`` and evaluate only that (``oxdna_device.cuh: bonded_pair / unbonded_pair``).
"""

from __future__ import annotations

import dataclasses as dc
import functools
from typing import Any

from mythos_b200.energy import base as je_base
from mythos_b200.energy.configuration import BaseConfiguration
from mythos_b200.energy.dna1 import terms as dna1
from mythos_b200.energy.dna2 import terms as dna2
from mythos_b200.energy.nucleotide import HybridNucleotide
from mythos_b200.energy.rna2 import terms as rna2
from mythos_b200.energy.rna2 import GEOMETRY_KEYS as _RNA_GEOMETRY_KEYS
from mythos_b200.energy.utils import default_configs_for


class _HybridConfiguration(BaseConfiguration):
    """Base of the NA1 configurations: prefixed copies of the per-bank parameters + shared ones."""

    BANKS: dict[str, type[BaseConfiguration]] = {}
    SHARED: tuple[str, ...] = ()
    REQUIRED_BANKS: tuple[str, ...] = ("dna", "rna")

    @classmethod
    def _bank_fields(cls, bank: str, *, required_only: bool) -> list[str]:
        sub = cls.BANKS[bank]
        names = sub.required_params if required_only else (*sub.required_params, *sub.optional_params)
        return [n for n in names if n not in cls.SHARED]

    def __init_subclass__(cls, **kw) -> None:
        super().__init_subclass__(**kw)
        if not cls.BANKS:
            return
        req = ["nt_type", *cls.SHARED]
        opt = []
        for bank in cls.BANKS:
            fields = [f"{bank}_{n}" for n in cls._bank_fields(bank, required_only=True)]
            (req if bank in cls.REQUIRED_BANKS else opt).extend(fields)
            opt += [f"{bank}_{n}" for n in cls.BANKS[bank].optional_params if n not in cls.SHARED and n not in ("pseq", "pseq_constraints")]
        cls.required_params = tuple(req)
        cls.optional_params = tuple(dict.fromkeys(opt))
        cls.dependent_params = tuple(f"{bank}_config" for bank in cls.BANKS)
        cls.non_optimizable_required_params = ("nt_type",)

    def init_params(self):
        out = {}
        for bank, sub in self.BANKS.items():
            kw = {n: getattr(self, f"{bank}_{n}") for n in self._bank_fields(bank, required_only=False) if f"{bank}_{n}" in self}
            kw.update({n: getattr(self, n) for n in self.SHARED})
            if any(kw.get(n) is None for n in sub.required_params):
                out[f"{bank}_config"] = None  # bank not configured (e.g. no hybrid parameters given)
                continue
            out[f"{bank}_config"] = sub(**kw).init_params()
        return self.replace(**out)


class FeneConfiguration(_HybridConfiguration):
    """na1/fene.py:19-81"""

    term = "fene"
    BANKS = {"dna": dna1.FeneConfiguration, "rna": dna1.FeneConfiguration}


class BondedExcludedVolumeConfiguration(_HybridConfiguration):
    """na1/bonded_excluded_volume.py"""

    term = "bonded_excluded_volume"
    BANKS = {"dna": dna1.BondedExcludedVolumeConfiguration, "rna": dna1.BondedExcludedVolumeConfiguration}


class StackingConfiguration(_HybridConfiguration):
    """na1/stacking.py:20-188 (DNA bank: dna1 configuration evaluated by dna2.Stacking; RNA bank: rna2)"""

    term = "stacking"
    BANKS = {"dna": dna1.StackingConfiguration, "rna": rna2.StackingConfiguration}
    SHARED = ("kt",)


class UnbondedExcludedVolumeConfiguration(_HybridConfiguration):
    """na1/unbonded_excluded_volume.py"""

    term = "unbonded_excluded_volume"
    BANKS = {"dna": dna1.UnbondedExcludedVolumeConfiguration, "rna": dna1.UnbondedExcludedVolumeConfiguration,
             "drh": dna1.UnbondedExcludedVolumeConfiguration}
    REQUIRED_BANKS = ("dna", "rna", "drh")


class HydrogenBondingConfiguration(_HybridConfiguration):
    """na1/hydrogen_bonding.py"""

    term = "hydrogen_bonding"
    BANKS = {"dna": dna1.HydrogenBondingConfiguration, "rna": dna1.HydrogenBondingConfiguration,
             "drh": dna1.HydrogenBondingConfiguration}
    REQUIRED_BANKS = ("dna", "rna", "drh")


class CrossStackingConfiguration(_HybridConfiguration):
    """na1/cross_stacking.py (RNA bank uses the rna2 form without theta4)"""

    term = "cross_stacking"
    BANKS = {"dna": dna1.CrossStackingConfiguration, "rna": rna2.CrossStackingConfiguration,
             "drh": dna1.CrossStackingConfiguration}
    REQUIRED_BANKS = ("dna", "rna", "drh")


class CoaxialStackingConfiguration(_HybridConfiguration):
    """na1/coaxial_stacking.py:160-240 (DNA bank: dna2 form; RNA and hybrid banks: dna1 form)"""

    term = "coaxial_stacking"
    BANKS = {"dna": dna2.CoaxialStackingConfiguration, "rna": dna1.CoaxialStackingConfiguration,
             "drh": dna1.CoaxialStackingConfiguration}
    REQUIRED_BANKS = ("dna", "rna", "drh")


class DebyeConfiguration(_HybridConfiguration):
    """na1/debye.py:17-95"""

    term = "debye"
    BANKS = {"dna": dna2.DebyeConfiguration, "rna": dna2.DebyeConfiguration, "drh": dna2.DebyeConfiguration}
    SHARED = ("half_charged_ends", "kt", "salt_conc")


@dc.dataclass(frozen=True, kw_only=True)
class _HybridTerm(je_base.BaseEnergyFunction):
    HYBRID = True


@dc.dataclass(frozen=True, kw_only=True)
class Fene(_HybridTerm):
    TERM = dna1.TERM_FENE


@dc.dataclass(frozen=True, kw_only=True)
class BondedExcludedVolume(_HybridTerm):
    TERM = dna1.TERM_BEXC


@dc.dataclass(frozen=True, kw_only=True)
class Stacking(_HybridTerm):
    TERM = dna1.TERM_STACK


@dc.dataclass(frozen=True, kw_only=True)
class UnbondedExcludedVolume(_HybridTerm):
    TERM = dna1.TERM_UEXC


@dc.dataclass(frozen=True, kw_only=True)
class HydrogenBonding(_HybridTerm):
    TERM = dna1.TERM_HB


@dc.dataclass(frozen=True, kw_only=True)
class CrossStacking(_HybridTerm):
    TERM = dna1.TERM_CROSS


@dc.dataclass(frozen=True, kw_only=True)
class CoaxialStacking(_HybridTerm):
    TERM = dna1.TERM_COAX


@dc.dataclass(frozen=True, kw_only=True)
class Debye(_HybridTerm):
    is_end: Any = None
    TERM = dna1.TERM_DEBYE

    def __post_init__(self, topology) -> None:
        super().__post_init__(topology)
        if topology is not None:
            object.__setattr__(self, "is_end", topology.is_end)
        if self.is_end is None:
            raise ValueError("is_end must be provided either through topology or directly.")

    def extra_topology(self) -> dict:
        return super().extra_topology() | {"is_end": self.is_end}


def default_transform_fn():
    """HybridNucleotide transform with the dna2 and rna2 default geometry (na1/tests/test_integration.py:107-133)."""
    gd = default_configs_for("dna2")[1]["geometry"]
    gr = default_configs_for("rna2")[1]["geometry"]
    kw = {f"dna_{k}": gd[k] for k in ("com_to_backbone_x", "com_to_backbone_y", "com_to_backbone_dna1", "com_to_hb", "com_to_stacking")}
    kw.update({f"rna_{k}": gr[v] for k, v in _RNA_GEOMETRY_KEYS.items()})
    return functools.partial(HybridNucleotide.from_rigid_body, **kw)


def default_params() -> dict[str, dict]:
    """Merged ``rna_`` / ``dna_`` / ``drh_`` parameter tables (na1/tests/test_integration.py:137-141)."""
    out: dict[str, dict] = {}
    for pre, model in (("rna_", "rna2"), ("dna_", "dna2"), ("drh_", "na1")):
        for term, vals in default_configs_for(model)[1].items():
            if term == "geometry":
                continue
            out.setdefault(term, {}).update({pre + k: v for k, v in vals.items()})
    return out


def default_energy_fns() -> list[type[je_base.BaseEnergyFunction]]:
    return [Fene, BondedExcludedVolume, Stacking, UnbondedExcludedVolume, HydrogenBonding, CrossStacking, CoaxialStacking, Debye]


def default_energy_configs(nt_type, kt=None, salt_conc=None, half_charged_ends=None, stack_nt_type=None) -> list[BaseConfiguration]:
    sim = default_configs_for("dna2")[0]
    p = default_params()
    kt = sim["kT"] if kt is None else kt
    shared = {
        "kt": kt,
        "salt_conc": sim["salt_conc"] if salt_conc is None else salt_conc,
        "half_charged_ends": bool(sim["half_charged_ends"]) if half_charged_ends is None else half_charged_ends,
    }
    nt = {"nt_type": nt_type}
    return [
        FeneConfiguration(**(p["fene"] | nt), params_to_optimize=BaseConfiguration.OPT_ALL),
        BondedExcludedVolumeConfiguration(**(p["bonded_excluded_volume"] | nt), params_to_optimize=BaseConfiguration.OPT_ALL),
        StackingConfiguration(**(p["stacking"] | {"nt_type": nt_type if stack_nt_type is None else stack_nt_type, "kt": kt}),
                              params_to_optimize=BaseConfiguration.OPT_ALL),
        UnbondedExcludedVolumeConfiguration(**(p["unbonded_excluded_volume"] | nt), params_to_optimize=BaseConfiguration.OPT_ALL),
        HydrogenBondingConfiguration(**(p["hydrogen_bonding"] | nt), params_to_optimize=BaseConfiguration.OPT_ALL),
        CrossStackingConfiguration(**(p["cross_stacking"] | nt), params_to_optimize=BaseConfiguration.OPT_ALL),
        CoaxialStackingConfiguration(**(p["coaxial_stacking"] | nt), params_to_optimize=BaseConfiguration.OPT_ALL),
        DebyeConfiguration(**(p["debye"] | nt | shared), params_to_optimize=BaseConfiguration.OPT_ALL),
    ]


def create_default_energy_fn(topology, displacement_fn=None, **config_kwargs) -> je_base.EnergyFunction:
    from mythos_b200.energy import DEFAULT_DISPLACEMENT

    return je_base.ComposedEnergyFunction.from_lists(
        energy_fns=default_energy_fns(),
        energy_configs=default_energy_configs(topology.nt_type, **config_kwargs),
        transform_fn=default_transform_fn(),
        displacement_fn=DEFAULT_DISPLACEMENT if displacement_fn is None else displacement_fn,
        topology=topology,
    )


__all__ = [
    "BondedExcludedVolume", "BondedExcludedVolumeConfiguration", "CoaxialStacking", "CoaxialStackingConfiguration",
    "CrossStacking", "CrossStackingConfiguration", "Debye", "DebyeConfiguration", "Fene", "FeneConfiguration",
    "HybridNucleotide", "HydrogenBonding", "HydrogenBondingConfiguration", "Stacking", "StackingConfiguration",
    "UnbondedExcludedVolume", "UnbondedExcludedVolumeConfiguration", "create_default_energy_fn", "default_energy_configs",
    "default_energy_fns", "default_params", "default_transform_fn",
]
