"""Pins the CPU oracle against the reference's own oxDNA-standalone golden energies.

Mirrors mythos/energy/{dna1,dna2,rna2,na1}/tests/test_integration.py: per-term energy of every frame,
divided by N, rounded to 6 dp, compared with split_energy.dat under the reference's tolerances, with the
periodic(20.0) displacement those tests hard-code.
"""

import numpy as np
import pytest
import torch

from oracle import oxdna_oracle as orc
from tests.golden_cases import ALL_CASES, TOL, load_case, stack_nt_type, theta_for


@pytest.mark.parametrize("name", ALL_CASES)
def test_terms_match_oxdna_golden(name):
    c = load_case(name)
    model = c["model"]
    params = orc.init_all(model, theta_for(c))
    n = c["center"].shape[1]
    got = []
    for f in range(c["center"].shape[0]):
        t = orc.energy_terms(
            model,
            c["center"][f],
            c["quat"][f],
            c["seq"],
            c["bonded"],
            c["pairs"],
            params,
            box=20.0,
            is_end=c["is_end"],
            nt_type=c["nt_type"],
            stack_nt_type=stack_nt_type(c),
        )
        got.append(t.numpy())
    got = np.around(np.stack(got) / n, 6)
    want = c["golden_terms_per_nt"]
    for k, term in enumerate(orc.TERMS):
        if model == "dna1" and term == "debye":
            continue
        atol = TOL[model][k]
        if "coax" in name and term != "coaxial_stacking":
            # the reference only checks the coaxial term on the coax fixtures; the others are extra coverage
            atol = max(atol, 1e-4)
        np.testing.assert_allclose(got[:, k], want[:, k], atol=atol, rtol=1e-7, err_msg=f"{name}:{term}")


@pytest.mark.parametrize("name", ["dna1_simple_helix", "dna2_simple_helix", "dna2_half_charged"])
def test_total_matches_energy_dat(name):
    # dna1/tests/test_integration.py:322-389 (rtol 1e-5, atol 1e-6 on split sum), dna2: atol 1e-3 vs energy.dat
    c = load_case(name)
    params = orc.init_all(c["model"], theta_for(c))
    n = c["center"].shape[1]
    tot = []
    for f in range(c["center"].shape[0]):
        t = orc.energy_terms(
            c["model"], c["center"][f], c["quat"][f], c["seq"], c["bonded"], c["pairs"], params, box=20.0, is_end=c["is_end"]
        )
        tot.append(float(t.sum()) / n)
    np.testing.assert_allclose(np.around(tot, 6), c["golden_potential_per_nt"], atol=1e-3)


def test_weights_and_neff_kat():
    # mythos/optimization/tests/test_objective.py:187-205: equal energies -> uniform weights, neff 1
    e = torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64)
    w, neff = orc.weights_and_neff(torch.tensor(1.0, dtype=torch.float64), e, e)
    np.testing.assert_allclose(w.numpy(), [1 / 3] * 3)
    np.testing.assert_allclose(float(neff), 1.0)
