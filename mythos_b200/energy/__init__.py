"""oxDNA-family energy functions (interface of ``mythos.energy``) over the fused sm_100a kernels."""

from mythos_b200 import space

DEFAULT_DISPLACEMENT = space.free()[0]

from mythos_b200.energy.base import (  # noqa: E402
    BaseEnergyFunction,
    ComposedEnergyFunction,
    EnergyFunction,
    QualifiedComposedEnergyFunction,
)
from mythos_b200.energy.configuration import BaseConfiguration  # noqa: E402

__all__ = [
    "DEFAULT_DISPLACEMENT",
    "BaseConfiguration",
    "BaseEnergyFunction",
    "ComposedEnergyFunction",
    "EnergyFunction",
    "QualifiedComposedEnergyFunction",
]
