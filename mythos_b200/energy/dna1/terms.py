"""oxDNA1 terms: configurations (independent -> dependent parameters) and term classes.

Same classes, parameter names and ``init_params`` formulas as ``mythos/energy/dna1/{fene,bonded_excluded_volume,
stacking,unbonded_excluded_volume,hydrogen_bonding,cross_stacking,coaxial_stacking}.py``.  The per-pair arithmetic
the reference keeps in ``pairwise_energies`` / ``dna1/interactions.py`` lives in the CUDA kernels
(``mythos_b200/csrc/oxdna_device.cuh``); a term class here only names its kernel term and functional form.
"""

from __future__ import annotations

import dataclasses as dc

import torch

from mythos_b200.energy import base as je_base
from mythos_b200.energy import base_smoothing_functions as bsf
from mythos_b200.energy.configuration import BaseConfiguration

TERM_FENE, TERM_BEXC, TERM_STACK, TERM_UEXC, TERM_HB, TERM_CROSS, TERM_COAX, TERM_DEBYE = range(8)

STACK_WEIGHTS_SA = torch.ones(4, 4, dtype=torch.float64)
HB_WEIGHTS_SA = torch.tensor([[0, 0, 0, 1], [0, 0, 1, 0], [0, 1, 0, 0], [1, 0, 0, 0]], dtype=torch.float64)


def _f4_fields(fam: str, ks: str | tuple) -> tuple[tuple[str, ...], tuple[str, ...]]:
    req, dep = [], []
    for k in ks:
        req += [f"a_{fam}_{k}", f"theta0_{fam}_{k}", f"delta_theta_star_{fam}_{k}"]
        dep += [f"b_{fam}_{k}", f"delta_theta_{fam}_{k}_c"]
    return tuple(req), tuple(dep)


def _stacked(cfg, names) -> torch.Tensor:
    vals = [bsf.as_t(getattr(cfg, nm)) for nm in names]
    return torch.stack([v if v.dim() == 0 else v.reshape(()) for v in vals])


def _f4_init(cfg, fam: str, ks) -> dict:
    """All f4 blocks of one term in ONE vectorised evaluation of the smoothing formulas (same arithmetic per element): the
    autograd graph the theta -> bank chain leaves behind is what the end-to-end DiffTRe step pays for on the host, and it
    is a third as long this way."""
    ks = list(ks)
    b, dc = bsf.get_f4_smoothing_params(
        _stacked(cfg, [f"a_{fam}_{k}" for k in ks]), _stacked(cfg, [f"theta0_{fam}_{k}" for k in ks]),
        _stacked(cfg, [f"delta_theta_star_{fam}_{k}" for k in ks]),
    )
    out = {}
    for k, bk, dk in zip(ks, b.unbind(0), dc.unbind(0)):
        out[f"b_{fam}_{k}"], out[f"delta_theta_{fam}_{k}_c"] = bk, dk
    return out


# ------------------------------------------------------------------------------------------------ FENE
class FeneConfiguration(BaseConfiguration):
    """dna1/fene.py:12-28"""

    term = "fene"
    required_params = ("eps_backbone", "r0_backbone", "delta_backbone", "fmax", "finf")

    def init_params(self) -> "FeneConfiguration":
        return self


@dc.dataclass(frozen=True, kw_only=True)
class Fene(je_base.BaseEnergyFunction):
    """FENE backbone spring (dna1/fene.py:31-61)."""

    TERM = TERM_FENE


# ------------------------------------------------------------------------------------------------ excluded volume
class BondedExcludedVolumeConfiguration(BaseConfiguration):
    """dna1/bonded_excluded_volume.py:13-76"""

    term = "bonded_excluded_volume"
    _sites = ("base", "back_base", "base_back")
    required_params = (
        "eps_exc", "dr_star_base", "sigma_base", "sigma_back_base", "sigma_base_back", "dr_star_back_base", "dr_star_base_back",
    )
    dependent_params = ("b_base", "dr_c_base", "b_back_base", "dr_c_back_base", "b_base_back", "dr_c_base_back")

    def init_params(self):
        b, rc = bsf.get_f3_smoothing_params(  # all sites in one vectorised evaluation (see _f4_init)
            _stacked(self, [f"dr_star_{s}" for s in self._sites]), _stacked(self, [f"sigma_{s}" for s in self._sites])
        )
        out = {}
        for s, bs, rs in zip(self._sites, b.unbind(0), rc.unbind(0)):
            out[f"b_{s}"], out[f"dr_c_{s}"] = bs, rs
        return self.replace(**out)


@dc.dataclass(frozen=True, kw_only=True)
class BondedExcludedVolume(je_base.BaseEnergyFunction):
    """dna1/bonded_excluded_volume.py:79-119"""

    TERM = TERM_BEXC


class UnbondedExcludedVolumeConfiguration(BondedExcludedVolumeConfiguration):
    """dna1/unbonded_excluded_volume.py:15-97"""

    term = "unbonded_excluded_volume"
    _sites = ("base", "back_base", "base_back", "backbone")
    required_params = (
        "eps_exc", "dr_star_base", "sigma_base", "dr_star_back_base", "sigma_back_base", "dr_star_base_back",
        "sigma_base_back", "dr_star_backbone", "sigma_backbone",
    )
    dependent_params = (
        "b_base", "dr_c_base", "b_back_base", "dr_c_back_base", "b_base_back", "dr_c_base_back", "b_backbone", "dr_c_backbone",
    )


@dc.dataclass(frozen=True, kw_only=True)
class UnbondedExcludedVolume(je_base.BaseEnergyFunction):
    """dna1/unbonded_excluded_volume.py:100-151"""

    TERM = TERM_UEXC


# ------------------------------------------------------------------------------------------------ stacking
_STACK_F4_REQ, _STACK_F4_DEP = _f4_fields("stack", "456")


class StackingConfiguration(BaseConfiguration):
    """dna1/stacking.py:29-183"""

    term = "stacking"
    _f4 = ("4", "5", "6")
    required_params = (
        "eps_stack_base", "eps_stack_kt_coeff", "dr_low_stack", "dr_high_stack", "a_stack", "dr0_stack", "dr_c_stack",
        "theta0_stack_4", "delta_theta_star_stack_4", "a_stack_4", "theta0_stack_5", "delta_theta_star_stack_5", "a_stack_5",
        "theta0_stack_6", "delta_theta_star_stack_6", "a_stack_6", "neg_cos_phi1_star_stack", "a_stack_1",
        "neg_cos_phi2_star_stack", "a_stack_2", "kt",
    )
    optional_params = ("pseq", "pseq_constraints", "ss_stack_weights")
    dependent_params = (
        "b_low_stack", "dr_c_low_stack", "b_high_stack", "dr_c_high_stack", *_STACK_F4_DEP,
        "b_neg_cos_phi1_stack", "neg_cos_phi1_c_stack", "b_neg_cos_phi2_stack", "neg_cos_phi2_c_stack", "eps_stack",
    )

    def _eps_stack(self) -> torch.Tensor:
        kt, coeff = bsf.as_t(self.kt), bsf.as_t(self.eps_stack_kt_coeff)
        if self.ss_stack_weights is None:
            return (bsf.as_t(self.eps_stack_base) + coeff * kt) * STACK_WEIGHTS_SA
        return bsf.as_t(self.ss_stack_weights) * (1.0 - coeff + kt * 9.0 * coeff)

    def init_params(self):
        if self.pseq is not None and self.pseq_constraints is None:
            raise ValueError("pseq_constraints must be provided when pseq is provided.")
        out = {"eps_stack": self._eps_stack()}
        out["b_low_stack"], out["dr_c_low_stack"], out["b_high_stack"], out["dr_c_high_stack"] = bsf.get_f1_smoothing_params(
            self.dr0_stack, self.a_stack, self.dr_c_stack, self.dr_low_stack, self.dr_high_stack
        )
        out.update(_f4_init(self, "stack", self._f4))
        b5, c5 = bsf.get_f5_smoothing_params(
            _stacked(self, ["a_stack_1", "a_stack_2"]), _stacked(self, ["neg_cos_phi1_star_stack", "neg_cos_phi2_star_stack"])
        )
        for k, bk, ck in zip(("1", "2"), b5.unbind(0), c5.unbind(0)):
            out[f"b_neg_cos_phi{k}_stack"], out[f"neg_cos_phi{k}_c_stack"] = bk, ck
        return self.replace(**out)


@dc.dataclass(frozen=True, kw_only=True)
class Stacking(je_base.BaseEnergyFunction):
    """dna1/stacking.py:186-293: f1(r_stack) f4(th4) f4(th5') f4(th6') f5(-cos phi1) f5(-cos phi2) x eps_stack[seq_i, seq_j]."""

    TERM = TERM_STACK
    FORM = {"stack_form": 0, "use_back_stack": 0}


# ------------------------------------------------------------------------------------------------ hydrogen bonding
_HB_F4_REQ, _HB_F4_DEP = _f4_fields("hb", "123478")


class HydrogenBondingConfiguration(BaseConfiguration):
    """dna1/hydrogen_bonding.py:28-223"""

    term = "hydrogen_bonding"
    required_params = ("eps_hb", "a_hb", "dr0_hb", "dr_c_hb", "dr_low_hb", "dr_high_hb", *_HB_F4_REQ)
    optional_params = ("ss_hb_weights", "pseq", "pseq_constraints")
    dependent_params = ("b_low_hb", "dr_c_low_hb", "b_high_hb", "dr_c_high_hb", *_HB_F4_DEP, "eps_hb_weights")

    def init_params(self):
        if self.pseq is not None and self.pseq_constraints is None:
            raise ValueError("pseq_constraints must be provided when pseq is provided.")
        out = {
            "eps_hb_weights": HB_WEIGHTS_SA * bsf.as_t(self.eps_hb) if self.ss_hb_weights is None else bsf.as_t(self.ss_hb_weights)
        }
        out["b_low_hb"], out["dr_c_low_hb"], out["b_high_hb"], out["dr_c_high_hb"] = bsf.get_f1_smoothing_params(
            self.dr0_hb, self.a_hb, self.dr_c_hb, self.dr_low_hb, self.dr_high_hb
        )
        out.update(_f4_init(self, "hb", "123478"))
        return self.replace(**out)


@dc.dataclass(frozen=True, kw_only=True)
class HydrogenBonding(je_base.BaseEnergyFunction):
    """dna1/hydrogen_bonding.py:226-340: f1(r_hb) prod f4(th1,2,3,4,7,8) x eps_hb_weights[seq_i, seq_j]."""

    TERM = TERM_HB


# ------------------------------------------------------------------------------------------------ cross stacking
_CR_F4_REQ, _CR_F4_DEP = _f4_fields("cross", "123478")


class CrossStackingConfiguration(BaseConfiguration):
    """dna1/cross_stacking.py:15-183"""

    term = "cross_stacking"
    _f4 = "123478"
    required_params = ("dr_low_cross", "dr_high_cross", "k_cross", "r0_cross", "dr_c_cross", *_CR_F4_REQ)
    dependent_params = ("b_low_cross", "dr_c_low_cross", "b_high_cross", "dr_c_high_cross", *_CR_F4_DEP)

    def init_params(self):
        out = {}
        out["b_low_cross"], out["dr_c_low_cross"], out["b_high_cross"], out["dr_c_high_cross"] = bsf.get_f2_smoothing_params(
            self.r0_cross, self.dr_c_cross, self.dr_low_cross, self.dr_high_cross
        )
        out.update(_f4_init(self, "cross", self._f4))
        return self.replace(**out)


@dc.dataclass(frozen=True, kw_only=True)
class CrossStacking(je_base.BaseEnergyFunction):
    """dna1/cross_stacking.py:186-271"""

    TERM = TERM_CROSS
    FORM = {"cross_form": 0}


# ------------------------------------------------------------------------------------------------ coaxial stacking
_CX_F4_REQ, _CX_F4_DEP = _f4_fields("coax", "4156")


class CoaxialStackingConfiguration(BaseConfiguration):
    """dna1/coaxial_stacking.py:15-172"""

    term = "coaxial_stacking"
    required_params = (
        "dr_low_coax", "dr_high_coax", "k_coax", "dr0_coax", "dr_c_coax", *_CX_F4_REQ,
        "cos_phi3_star_coax", "a_coax_3p", "cos_phi4_star_coax", "a_coax_4p",
    )
    dependent_params = (
        "b_low_coax", "dr_c_low_coax", "b_high_coax", "dr_c_high_coax", *_CX_F4_DEP,
        "b_cos_phi3_coax", "cos_phi3_c_coax", "b_cos_phi4_coax", "cos_phi4_c_coax",
    )

    def init_params(self):
        out = {}
        out["b_low_coax"], out["dr_c_low_coax"], out["b_high_coax"], out["dr_c_high_coax"] = bsf.get_f2_smoothing_params(
            self.dr0_coax, self.dr_c_coax, self.dr_low_coax, self.dr_high_coax
        )
        out.update(_f4_init(self, "coax", "4156"))
        out["b_cos_phi3_coax"], out["cos_phi3_c_coax"] = bsf.get_f5_smoothing_params(self.a_coax_3p, self.cos_phi3_star_coax)
        out["b_cos_phi4_coax"], out["cos_phi4_c_coax"] = bsf.get_f5_smoothing_params(self.a_coax_4p, self.cos_phi4_star_coax)
        return self.replace(**out)


@dc.dataclass(frozen=True, kw_only=True)
class CoaxialStacking(je_base.BaseEnergyFunction):
    """dna1/coaxial_stacking.py:175-265"""

    TERM = TERM_COAX
    FORM = {"coax_form": 0}
