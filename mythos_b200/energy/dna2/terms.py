"""oxDNA2-specific terms: stacking on the oxDNA1 backbone site, f6 coaxial stacking, Debye-Hueckel.

Same classes / parameter names / ``init_params`` as ``mythos/energy/dna2/{stacking,coaxial_stacking,debye}.py``.
"""

from __future__ import annotations

import dataclasses as dc
from typing import Any

import torch

from mythos_b200.energy import base as je_base
from mythos_b200.energy import base_smoothing_functions as bsf
from mythos_b200.energy.configuration import BaseConfiguration
from mythos_b200.energy.dna1 import terms as dna1


@dc.dataclass(frozen=True, kw_only=True)
class Stacking(dna1.Stacking):
    """dna2/stacking.py:13-44: oxDNA1 stacking evaluated with ``back_sites_dna1`` for the cos(phi) factors."""

    FORM = {"stack_form": 0, "use_back_stack": 1}


_CX_F4_REQ, _CX_F4_DEP = dna1._f4_fields("coax", "4156")


class CoaxialStackingConfiguration(BaseConfiguration):
    """dna2/coaxial_stacking.py:15-130"""

    term = "coaxial_stacking"
    required_params = ("dr_low_coax", "dr_high_coax", "k_coax", "dr0_coax", "dr_c_coax", *_CX_F4_REQ, "a_coax_1_f6", "b_coax_1_f6")
    dependent_params = ("b_low_coax", "dr_c_low_coax", "b_high_coax", "dr_c_high_coax", *_CX_F4_DEP)

    def init_params(self):
        out = {}
        out["b_low_coax"], out["dr_c_low_coax"], out["b_high_coax"], out["dr_c_high_coax"] = bsf.get_f2_smoothing_params(
            self.dr0_coax, self.dr_c_coax, self.dr_low_coax, self.dr_high_coax
        )
        out.update(dna1._f4_init(self, "coax", "4156"))
        return self.replace(**out)


@dc.dataclass(frozen=True, kw_only=True)
class CoaxialStacking(je_base.BaseEnergyFunction):
    """dna2/coaxial_stacking.py:133-206: f2 f4(th4) [f4(th1)+f6(th1)] [f4(th5)+f4(pi-th5)] [f4(th6)+f4(pi-th6)]."""

    TERM = dna1.TERM_COAX
    FORM = {"coax_form": 1}


class DebyeConfiguration(BaseConfiguration):
    """dna2/debye.py:15-65"""

    term = "debye"
    required_params = ("q_eff", "lambda_factor", "prefactor_coeff", "kt", "salt_conc", "half_charged_ends")
    dependent_params = ("lambda_", "kappa", "r_high", "prefactor", "smoothing_coeff", "r_cut")

    def init_params(self):
        t = bsf.as_t
        lam = t(self.lambda_factor) * torch.sqrt(t(self.kt) / 0.1) / torch.sqrt(t(self.salt_conc))
        kappa = 1.0 / lam
        r_high = 3 * lam
        pref = t(self.prefactor_coeff) * t(self.q_eff) ** 2
        smoothing = -(torch.exp(-r_high / lam) * pref * pref * (r_high + lam) * (r_high + lam)) / (
            -4.0 * r_high * r_high * r_high * lam * lam * pref
        )
        r_cut = r_high * (pref * r_high + 3.0 * pref * lam) / (pref * (r_high + lam))
        return self.replace(lambda_=lam, kappa=kappa, r_high=r_high, prefactor=pref, smoothing_coeff=smoothing, r_cut=r_cut)


@dc.dataclass(frozen=True, kw_only=True)
class Debye(je_base.BaseEnergyFunction):
    """dna2/debye.py:68-115: exp(-kappa r) A / r on the backbone sites, quadratic smoothing, half-charged strand ends."""

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
