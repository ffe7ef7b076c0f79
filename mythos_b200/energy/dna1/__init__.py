"""oxDNA1 energy model (interface of ``mythos.energy.dna1``, ``mythos/energy/dna1/__init__.py:22-123``)."""

import functools
from types import MappingProxyType

from mythos_b200.energy import DEFAULT_DISPLACEMENT
from mythos_b200.energy.base import BaseEnergyFunction, ComposedEnergyFunction, EnergyFunction
from mythos_b200.energy.configuration import BaseConfiguration
from mythos_b200.energy.dna1.terms import (
    BondedExcludedVolume,
    BondedExcludedVolumeConfiguration,
    CoaxialStacking,
    CoaxialStackingConfiguration,
    CrossStacking,
    CrossStackingConfiguration,
    Fene,
    FeneConfiguration,
    HydrogenBonding,
    HydrogenBondingConfiguration,
    Stacking,
    StackingConfiguration,
    UnbondedExcludedVolume,
    UnbondedExcludedVolumeConfiguration,
)
from mythos_b200.energy.nucleotide import Dna1Nucleotide as Nucleotide
from mythos_b200.energy.utils import default_configs_for


def default_configs():
    """(simulation, energy) default constants of oxDNA1."""
    return default_configs_for("dna1")


def default_energy_configs(overrides: dict = MappingProxyType({}), opts: dict = MappingProxyType({})) -> list[BaseConfiguration]:
    sim, cfg = default_configs()

    def get_param(x: str) -> dict:
        return cfg[x] | overrides.get(x, {})

    def get_opts(x: str, defaults: tuple = BaseConfiguration.OPT_ALL) -> tuple:
        return opts.get(x, defaults)

    stacking_opts = tuple(set(cfg["stacking"].keys()) - {"kT", "ss_stack_weights"})
    return [
        FeneConfiguration.from_dict(get_param("fene"), get_opts("fene")),
        BondedExcludedVolumeConfiguration.from_dict(get_param("bonded_excluded_volume"), get_opts("bonded_excluded_volume")),
        StackingConfiguration.from_dict(get_param("stacking") | {"kt": overrides.get("kT", sim["kT"])}, get_opts("stacking", stacking_opts)),
        UnbondedExcludedVolumeConfiguration.from_dict(get_param("unbonded_excluded_volume"), get_opts("unbonded_excluded_volume")),
        HydrogenBondingConfiguration.from_dict(get_param("hydrogen_bonding"), get_opts("hydrogen_bonding")),
        CrossStackingConfiguration.from_dict(get_param("cross_stacking"), get_opts("cross_stacking")),
        CoaxialStackingConfiguration.from_dict(get_param("coaxial_stacking"), get_opts("coaxial_stacking")),
    ]


def default_energy_fns() -> list[type[BaseEnergyFunction]]:
    return [Fene, BondedExcludedVolume, Stacking, UnbondedExcludedVolume, HydrogenBonding, CrossStacking, CoaxialStacking]


def default_transform_fn():
    g = default_configs()[1]["geometry"]
    return functools.partial(
        Nucleotide.from_rigid_body,
        com_to_backbone=g["com_to_backbone"],
        com_to_hb=g["com_to_hb"],
        com_to_stacking=g["com_to_stacking"],
    )


def create_default_energy_fn(topology, displacement_fn=DEFAULT_DISPLACEMENT) -> EnergyFunction:
    """The default oxDNA1 composed energy function for a topology (``dna1/__init__.py:88-102``)."""
    return ComposedEnergyFunction.from_lists(
        energy_fns=default_energy_fns(),
        energy_configs=default_energy_configs(),
        transform_fn=default_transform_fn(),
        displacement_fn=displacement_fn,
        topology=topology,
    )


__all__ = [
    "BondedExcludedVolume", "BondedExcludedVolumeConfiguration", "CoaxialStacking", "CoaxialStackingConfiguration",
    "CrossStacking", "CrossStackingConfiguration", "Fene", "FeneConfiguration", "HydrogenBonding",
    "HydrogenBondingConfiguration", "Nucleotide", "Stacking", "StackingConfiguration", "UnbondedExcludedVolume",
    "UnbondedExcludedVolumeConfiguration", "create_default_energy_fn", "default_configs", "default_energy_configs",
    "default_energy_fns", "default_transform_fn",
]
