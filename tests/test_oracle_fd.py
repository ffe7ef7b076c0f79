"""Self-validation of the ORACLE's derivatives by central finite differences (oracle/oxdna_oracle.py's header promises it).

The reference never differentiates a real energy term in its tests (derivatives unpinned), so the oracle's
torch autograd is what the CUDA gradients are judged against; before it may judge anything it is checked here against
central differences of its OWN energies: dE/dcenter, dE/dquat on randomly chosen coordinates, and dE/dtheta for
independent parameters that reach the kernel bank through every kind of smoothing solver (f1..f5, Debye).
"""

import numpy as np
import pytest
import torch

from oracle import oxdna_oracle as orc
from tests.golden_cases import load_case, stack_nt_type, theta_for

CASES = ["dna1_simple_helix", "dna2_half_charged", "dna2_simple_coax", "rna2_helix_12bp", "na1_helix_dna_rna", "na1_coax_rna"]
THETA = {
    "dna1": [("fene", "eps_backbone"), ("fene", "r0_backbone"), ("unbonded_excluded_volume", "sigma_backbone"), ("stacking", "a_stack"),
             ("stacking", "theta0_stack_5"), ("stacking", "neg_cos_phi1_star_stack"), ("hydrogen_bonding", "dr0_hb"),
             ("hydrogen_bonding", "a_hb_7"), ("cross_stacking", "r0_cross"), ("coaxial_stacking", "k_coax")],
    "dna2": [("fene", "eps_backbone"), ("bonded_excluded_volume", "dr_star_base"), ("stacking", "dr_c_stack"), ("stacking", "kt"),
             ("hydrogen_bonding", "a_hb"), ("hydrogen_bonding", "delta_theta_star_hb_4"), ("cross_stacking", "theta0_cross_8"),
             ("coaxial_stacking", "a_coax_1_f6"), ("coaxial_stacking", "dr0_coax"), ("debye", "q_eff"), ("debye", "lambda_factor")],
    "rna2": [("fene", "eps_backbone"), ("stacking", "a_stack_9"), ("stacking", "theta0_stack_10"), ("hydrogen_bonding", "eps_hb"),
             ("cross_stacking", "a_cross_7"), ("coaxial_stacking", "k_coax"), ("debye", "q_eff")],
    "na1": [("fene", "dna_eps_backbone"), ("fene", "rna_r0_backbone"), ("stacking", "rna_a_stack_9"), ("stacking", "dna_a_stack_4"),
            ("hydrogen_bonding", "drh_eps_hb"), ("hydrogen_bonding", "rna_eps_hb"), ("cross_stacking", "drh_k_cross"),
            ("coaxial_stacking", "rna_k_coax"), ("debye", "drh_q_eff"), ("unbonded_excluded_volume", "drh_sigma_backbone")],
}
FRAME = 23


def _energy(c, center, quat, theta, cot):
    params = orc.init_all(c["model"], theta)
    t = orc.energy_terms(c["model"], center, quat, c["seq"], c["bonded"], c["pairs"], params, box=20.0, is_end=c["is_end"],
                         nt_type=c["nt_type"], stack_nt_type=stack_nt_type(c))
    return (t * cot).sum()


@pytest.mark.parametrize("name", CASES)
def test_autograd_matches_central_differences(name):
    c = load_case(name)
    rng = np.random.default_rng(11)
    cot = torch.tensor(rng.uniform(0.5, 1.5, size=8))
    theta = theta_for(c)
    leaves = {}
    for term, nm in THETA[c["model"]]:
        leaves[(term, nm)] = torch.tensor(float(theta[term][nm]), dtype=torch.float64, requires_grad=True)
        for t2 in theta:  # shared names (kt, eps_exc, ...) are one parameter in the global namespace
            if nm in theta[t2]:
                theta[t2][nm] = leaves[(term, nm)]
    center = torch.tensor(c["center"][FRAME], requires_grad=True)
    quat = torch.tensor(c["quat"][FRAME], requires_grad=True)
    _energy(c, center, quat, theta, cot).backward()

    def e_at(dc=None, dq=None, th=None):
        with torch.no_grad():
            cc = center.detach().clone()
            qq = quat.detach().clone()
            if dc is not None:
                cc[dc[0], dc[1]] += dc[2]
            if dq is not None:
                qq[dq[0], dq[1]] += dq[2]
            t2 = {term: dict(vals) for term, vals in theta_for(c).items()}
            if th is not None:
                for term in t2:
                    if th[0][1] in t2[term]:
                        t2[term][th[0][1]] = float(t2[term][th[0][1]]) + th[1]
            return float(_energy(c, cc, qq, t2, cot))

    n = center.shape[0]
    h = 1e-6
    for _ in range(6):
        i, d = int(rng.integers(n)), int(rng.integers(3))
        fd = (e_at(dc=(i, d, h)) - e_at(dc=(i, d, -h))) / (2 * h)
        assert np.isclose(float(center.grad[i, d]), fd, rtol=2e-6, atol=2e-7), ("center", i, d, float(center.grad[i, d]), fd)
        i, d = int(rng.integers(n)), int(rng.integers(4))
        fd = (e_at(dq=(i, d, h)) - e_at(dq=(i, d, -h))) / (2 * h)
        assert np.isclose(float(quat.grad[i, d]), fd, rtol=2e-6, atol=2e-7), ("quat", i, d, float(quat.grad[i, d]), fd)
    for key, leaf in leaves.items():
        hh = 1e-6 * max(1.0, abs(float(leaf.detach())))
        fd = (e_at(th=(key, hh)) - e_at(th=(key, -hh))) / (2 * hh)
        got = 0.0 if leaf.grad is None else float(leaf.grad)
        assert np.isclose(got, fd, rtol=5e-6, atol=5e-7), (key, got, fd)
