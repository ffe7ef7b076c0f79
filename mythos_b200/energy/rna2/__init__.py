"""oxRNA2 energy model (interface of ``mythos.energy.rna2``) plus a default factory assembled the way the
reference's RNA2 integration tests assemble it (``mythos/energy/rna2/tests/test_integration.py:51-82``)."""

import functools
from types import MappingProxyType

from mythos_b200.energy import DEFAULT_DISPLACEMENT
from mythos_b200.energy.base import BaseEnergyFunction, ComposedEnergyFunction, EnergyFunction
from mythos_b200.energy.configuration import BaseConfiguration
from mythos_b200.energy.dna1 import terms as dna1
from mythos_b200.energy.dna2 import terms as dna2
from mythos_b200.energy.nucleotide import Rna2Nucleotide as Nucleotide
from mythos_b200.energy.rna2.terms import CrossStacking, CrossStackingConfiguration, Stacking, StackingConfiguration
from mythos_b200.energy.utils import default_configs_for

GEOMETRY_KEYS = {  # transform_fn keyword -> [geometry] key of the RNA2 constants
    "com_to_backbone_x": "pos_back_a1", "com_to_backbone_y": "pos_back_a3", "com_to_hb": "pos_base",
    "com_to_stacking": "pos_stack", "p3_x": "p3_x", "p3_y": "p3_y", "p3_z": "p3_z", "p5_x": "p5_x", "p5_y": "p5_y",
    "p5_z": "p5_z", "pos_stack_3_a1": "pos_stack_3_a1", "pos_stack_3_a2": "pos_stack_3_a2",
    "pos_stack_5_a1": "pos_stack_5_a1", "pos_stack_5_a2": "pos_stack_5_a2",
}


def default_configs():
    return default_configs_for("rna2")


def default_transform_fn():
    g = default_configs()[1]["geometry"]
    return functools.partial(Nucleotide.from_rigid_body, **{k: g[v] for k, v in GEOMETRY_KEYS.items()})


def default_energy_fns() -> list[type[BaseEnergyFunction]]:
    return [dna1.Fene, dna1.BondedExcludedVolume, Stacking, dna1.UnbondedExcludedVolume, dna1.HydrogenBonding,
            CrossStacking, dna1.CoaxialStacking, dna2.Debye]


def default_energy_configs(overrides: dict = MappingProxyType({}), opts: dict = MappingProxyType({})) -> list[BaseConfiguration]:
    sim, cfg = default_configs()

    def get_param(x: str) -> dict:
        return cfg[x] | overrides.get(x, {})

    def get_opts(x: str, defaults: tuple = BaseConfiguration.OPT_ALL) -> tuple:
        return opts.get(x, defaults)

    kt = overrides.get("kT", sim["kT"])
    debye_over = {
        "kt": kt,
        "salt_conc": overrides.get("salt_conc", sim["salt_conc"]),
        "half_charged_ends": overrides.get("half_charged_ends", bool(sim["half_charged_ends"])),
    }
    stacking_opts = tuple(set(cfg["stacking"].keys()) - {"kT", "ss_stack_weights"})
    debye_opts = tuple(set(cfg["debye"].keys()) - {"kT", "salt_conc"})
    return [
        dna1.FeneConfiguration.from_dict(get_param("fene"), get_opts("fene")),
        dna1.BondedExcludedVolumeConfiguration.from_dict(get_param("bonded_excluded_volume"), get_opts("bonded_excluded_volume")),
        StackingConfiguration.from_dict(get_param("stacking") | {"kt": kt}, get_opts("stacking", stacking_opts)),
        dna1.UnbondedExcludedVolumeConfiguration.from_dict(get_param("unbonded_excluded_volume"), get_opts("unbonded_excluded_volume")),
        dna1.HydrogenBondingConfiguration.from_dict(get_param("hydrogen_bonding"), get_opts("hydrogen_bonding")),
        CrossStackingConfiguration.from_dict(get_param("cross_stacking"), get_opts("cross_stacking")),
        dna1.CoaxialStackingConfiguration.from_dict(get_param("coaxial_stacking"), get_opts("coaxial_stacking")),
        dna2.DebyeConfiguration.from_dict(get_param("debye") | debye_over, get_opts("debye", debye_opts)),
    ]


def create_default_energy_fn(topology, displacement_fn=DEFAULT_DISPLACEMENT, **config_kwargs) -> EnergyFunction:
    return ComposedEnergyFunction.from_lists(
        energy_fns=default_energy_fns(),
        energy_configs=default_energy_configs(**config_kwargs),
        transform_fn=default_transform_fn(),
        displacement_fn=displacement_fn,
        topology=topology,
    )


__all__ = ["CrossStacking", "CrossStackingConfiguration", "Nucleotide", "Stacking", "StackingConfiguration",
           "create_default_energy_fn", "default_configs", "default_energy_configs", "default_energy_fns", "default_transform_fn"]
