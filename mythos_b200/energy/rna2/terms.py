"""oxRNA2-specific terms (``mythos/energy/rna2/{stacking,cross_stacking}.py``); the rest is reused from dna1/dna2."""

from __future__ import annotations

import dataclasses as dc

from mythos_b200.energy import base as je_base
from mythos_b200.energy import base_smoothing_functions as bsf
from mythos_b200.energy.dna1 import terms as dna1

_ST_REQ, _ST_DEP = dna1._f4_fields("stack", ("5", "6", "9", "10"))


class StackingConfiguration(dna1.StackingConfiguration):
    """rna2/stacking.py:21-175: theta5, theta6, theta9, theta10 factors; eps_stack = ss * (1 + kT * coeff) when sequence-specific."""

    _f4 = ("5", "6", "9", "10")
    required_params = (
        "eps_stack_base", "eps_stack_kt_coeff", "dr_low_stack", "dr_high_stack", "a_stack", "dr0_stack", "dr_c_stack",
        "theta0_stack_5", "delta_theta_star_stack_5", "a_stack_5", "theta0_stack_6", "delta_theta_star_stack_6", "a_stack_6",
        "theta0_stack_9", "delta_theta_star_stack_9", "a_stack_9", "theta0_stack_10", "delta_theta_star_stack_10", "a_stack_10",
        "neg_cos_phi1_star_stack", "a_stack_1", "neg_cos_phi2_star_stack", "a_stack_2", "kt",
    )
    dependent_params = (
        "b_low_stack", "dr_c_low_stack", "b_high_stack", "dr_c_high_stack", *_ST_DEP,
        "b_neg_cos_phi1_stack", "neg_cos_phi1_c_stack", "b_neg_cos_phi2_stack", "neg_cos_phi2_c_stack", "eps_stack",
    )

    def _eps_stack(self):
        kt, coeff = bsf.as_t(self.kt), bsf.as_t(self.eps_stack_kt_coeff)
        if self.ss_stack_weights is not None:
            return bsf.as_t(self.ss_stack_weights) * (1.0 + kt * coeff)
        return (bsf.as_t(self.eps_stack_base) + coeff * kt) * dna1.STACK_WEIGHTS_SA


@dc.dataclass(frozen=True, kw_only=True)
class Stacking(dna1.Stacking):
    """rna2/stacking.py:178-293: r(stack5_i, stack3_j), theta9/theta10 from the p3/p5 backbone directions."""

    FORM = {"stack_form": 1, "use_back_stack": 0}


_CR_REQ, _CR_DEP = dna1._f4_fields("cross", "12378")


class CrossStackingConfiguration(dna1.CrossStackingConfiguration):
    """rna2/cross_stacking.py:15-148 (no theta4 factor)"""

    _f4 = "12378"
    required_params = ("dr_low_cross", "dr_high_cross", "k_cross", "r0_cross", "dr_c_cross", *_CR_REQ)
    dependent_params = ("b_low_cross", "dr_c_low_cross", "b_high_cross", "dr_c_high_cross", *_CR_DEP)


@dc.dataclass(frozen=True, kw_only=True)
class CrossStacking(je_base.BaseEnergyFunction):
    """rna2/cross_stacking.py:151-228"""

    TERM = dna1.TERM_CROSS
    FORM = {"cross_form": 1}
