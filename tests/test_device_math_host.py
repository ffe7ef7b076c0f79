"""CPU check of the device math header against the oracle (values and all three gradient families).

The exact expressions the CUDA kernels run (mythos_b200/csrc/oxdna_device.cuh) are compiled for the host by
tests/hostcheck.py and compared with torch autograd on the oracle restatement: per-term energies, dE/dcenter,
dE/dquat, and dE/d(kernel-level parameters), for every model family, with a random per-term cotangent.
Tolerance: 1e-6 relative (the north star's float64 bar) on quantities scaled by their max magnitude.
"""

import numpy as np
import pytest
import torch

from mythos_b200 import _lib
from mythos_b200.energy import model as kmodel
from oracle import oxdna_oracle as orc
from tests import hostcheck
from tests.golden_cases import load_case, stack_nt_type, theta_for
from tests.product_cases import energy_fn_of

CASES = ["dna1_simple_helix", "dna1_seq_dep", "dna1_simple_coax", "dna2_half_charged", "dna2_simple_coax",
         "rna2_helix_12bp", "rna2_simple_coax", "na1_helix_dna_rna", "na1_helix_rna_dna", "na1_coax_dna", "na1_coax_rna"]


def _leafify(p):
    if isinstance(p, dict):
        return {k: _leafify(v) for k, v in p.items()}
    if isinstance(p, bool) or p is None:
        return p
    if isinstance(p, (float, int)):
        return torch.tensor(float(p), dtype=torch.float64, requires_grad=True)
    if isinstance(p, torch.Tensor) and p.dtype == torch.float64:
        return p.detach().clone().requires_grad_(True)
    return p


def _oracle_param_grads(model, params_leaf, n_banks):
    """Scatter the oracle's parameter gradients into the kernel bank order."""
    P = _lib.param_count()
    out = np.zeros(n_banks * P)
    layout = kmodel.bank_layout()
    banks = [("dna", 0), ("rna", 1), ("drh", 2)] if model == "na1" else [(None, 0)]
    for bname, b in banks:
        tree = params_leaf[bname] if bname else params_leaf
        for term, entries in layout.items():
            if term not in tree:
                continue
            for idx, field, sub in entries:
                v = tree[term].get(field)
                if not isinstance(v, torch.Tensor) or v.grad is None:
                    continue
                g = v.grad
                out[b * P + idx] = float(g[sub]) if sub is not None else float(g)
    return out


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("frame", [0, 57])
def test_values_and_gradients_match_oracle(name, frame):
    c = load_case(name)
    model = c["model"]
    efn = energy_fn_of(c)
    plan = kmodel.plan_for(efn.energy_fns)
    pvec = plan.params_vector().detach().numpy()
    rng = np.random.default_rng(7)
    cot = rng.uniform(0.5, 1.5, size=8)
    snt = stack_nt_type(c)

    terms, d_center, d_quat, d_params = hostcheck.evaluate(
        plan.model, c["center"][frame], c["quat"][frame], c["seq"], c["bonded"], c["pairs"], pvec, cot=cot,
        nt_type=c["nt_type"], nt_type_stack=snt, is_end=c["is_end"],
    )

    params = _leafify(orc.init_all(model, theta_for(c)))
    center = torch.tensor(c["center"][frame], requires_grad=True)
    quat = torch.tensor(c["quat"][frame], requires_grad=True)
    t = orc.energy_terms(model, center, quat, c["seq"], c["bonded"], c["pairs"], params, box=20.0, is_end=c["is_end"],
                         nt_type=c["nt_type"], stack_nt_type=snt)
    (t * torch.as_tensor(cot)).sum().backward()

    np.testing.assert_allclose(terms, t.detach().numpy(), rtol=1e-9, atol=1e-10)
    for got, want in ((d_center, center.grad.numpy()), (d_quat, quat.grad.numpy())):
        scale = np.abs(want).max()
        np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-7 * scale)
    want_p = _oracle_param_grads(model, params, plan.model.n_banks)
    # table entries that are exactly zero are structural zeros for the kernels (HB_WEIGHTS_SA, hydrogen_bonding.py:18-25)
    names = _lib.param_names()
    P = _lib.param_count()
    for b in range(plan.model.n_banks):
        for i, nm in enumerate(names):
            if "[" in nm and pvec[b * P + i] == 0.0:
                want_p[b * P + i] = 0.0
    scale = max(np.abs(want_p).max(), 1.0)
    bad = np.where(~np.isclose(d_params, want_p, rtol=1e-6, atol=1e-9 * scale))[0]
    assert bad.size == 0, [(int(i) // P, names[int(i) % P] if int(i) % P < len(names) else "pad", d_params[i], want_p[i]) for i in bad[:10]]


def test_float32_matches_float64_within_1e4():
    c = load_case("dna2_half_charged")
    efn = energy_fn_of(c)
    plan = kmodel.plan_for(efn.energy_fns)
    pvec = plan.params_vector().detach().numpy()
    args = (plan.model, c["center"][3], c["quat"][3], c["seq"], c["bonded"], c["pairs"], pvec)
    kw = dict(nt_type=c["nt_type"], is_end=c["is_end"])
    t64, dc64, dq64, dp64 = hostcheck.evaluate(*args, **kw)
    t32, dc32, dq32, dp32 = hostcheck.evaluate(*args, use_f32=True, **kw)
    np.testing.assert_allclose(t32, t64, rtol=1e-4, atol=1e-4 * np.abs(t64).max())
    np.testing.assert_allclose(dc32, dc64, rtol=1e-4, atol=1e-4 * np.abs(dc64).max())
    np.testing.assert_allclose(dq32, dq64, rtol=1e-4, atol=1e-4 * np.abs(dq64).max())
