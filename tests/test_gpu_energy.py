"""GPU parity tests of the energy / force / theta-gradient kernels, through the product API and the C-ABI.

Mirrors the reference's integration tests (mythos/energy/{dna1,dna2,rna2,na1}/tests/test_integration.py): per-term
energies of all 100 golden frames against oxDNA's split_energy.dat with the reference's tolerances -- here the
100 frames go through ``map`` in one batched launch -- plus what the reference never tests: forces, dE/dquat and
dE/dtheta against the oracle's autograd (1e-6 relative in float64, 1e-4 in float32, the north star's bars).
"""

import numpy as np
import pytest
import torch

from mythos_b200.rigid_body import Quaternion, RigidBody
from oracle import oxdna_oracle as orc
from tests.golden_cases import ALL_CASES, TOL, load_case, stack_nt_type, theta_for
from tests.product_cases import energy_fn_of

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def body_of(case, frames=slice(None), dtype=torch.float64, requires_grad=False):
    c = torch.tensor(case["center"][frames], dtype=dtype, device=DEV, requires_grad=requires_grad)
    q = torch.tensor(case["quat"][frames], dtype=dtype, device=DEV, requires_grad=requires_grad)
    return RigidBody(c, Quaternion(q))


@pytest.mark.parametrize("name", ALL_CASES)
def test_golden_terms(name):
    c = load_case(name)
    efn = energy_fn_of(c)
    n = c["center"].shape[1]
    terms = efn.compute_terms_frames(body_of(c)).cpu().numpy()  # (100, T)
    got = np.around(terms / n, 6)
    want = c["golden_terms_per_nt"]
    for k in range(terms.shape[1]):
        atol = TOL[c["model"]][k]
        if "coax" in name and k != 6:
            atol = max(atol, 1e-4)
        np.testing.assert_allclose(got[:, k], want[:, k], atol=atol, rtol=1e-7, err_msg=f"{name}: term {k}")
    # map == sum of terms (weights are ones), single body == first frame
    e_map = efn.map(body_of(c)).cpu().numpy()
    np.testing.assert_allclose(e_map, terms.sum(1), rtol=1e-12)
    e0 = efn(body_of(c, 0))
    np.testing.assert_allclose(float(e0), e_map[0], rtol=1e-12)


@pytest.mark.parametrize("name", ALL_CASES)
def test_terms_match_oracle_f64(name):
    c = load_case(name)
    efn = energy_fn_of(c)
    params = orc.init_all(c["model"], theta_for(c))
    got = efn.compute_terms_frames(body_of(c, slice(0, 10))).cpu().numpy()
    for f in range(10):
        want = orc.energy_terms(c["model"], c["center"][f], c["quat"][f], c["seq"], c["bonded"], c["pairs"], params, box=20.0,
                                is_end=c["is_end"], nt_type=c["nt_type"], stack_nt_type=stack_nt_type(c)).numpy()
        want = want[: got.shape[1]]
        np.testing.assert_allclose(got[f], want, rtol=1e-9, atol=1e-10)


GRAD_CASES = ["dna1_simple_helix", "dna1_seq_dep", "dna2_half_charged", "dna2_simple_coax", "rna2_helix_12bp",
              "na1_helix_dna_rna", "na1_coax_rna"]
THETA = {
    "dna1": ["eps_backbone", "r0_backbone", "eps_exc", "sigma_base", "dr_star_backbone", "a_stack", "theta0_stack_5",
             "a_stack_1", "kt", "eps_hb", "a_hb_7", "theta0_hb_4", "k_cross", "a_cross_4", "dr0_coax", "a_coax_3p"],
    "dna2": ["eps_backbone", "eps_exc", "sigma_backbone", "a_stack", "delta_theta_star_stack_4", "kt", "a_hb", "dr_c_hb",
             "k_cross", "theta0_cross_8", "k_coax", "a_coax_1_f6", "q_eff", "lambda_factor", "salt_conc"],
    "rna2": ["eps_backbone", "eps_exc", "a_stack", "a_stack_9", "theta0_stack_10", "kt", "eps_hb", "k_cross", "a_cross_7",
             "k_coax", "q_eff"],
    "na1": ["dna_eps_backbone", "rna_r0_backbone", "rna_a_stack_9", "dna_a_stack_4", "kt", "drh_eps_hb", "rna_eps_hb",
            "drh_k_cross", "dna_k_cross", "rna_k_coax", "drh_q_eff", "dna_sigma_backbone", "drh_sigma_backbone"],
}


def _oracle_theta_grads(c, names, frame, cot):
    """d(sum_t cot_t E_t)/d theta for the named independent parameters (global namespace), by autograd on the oracle."""
    th = theta_for(c)
    leaves = {}
    for nm in names:
        val = None
        for term in th:
            if nm in th[term]:
                val = th[term][nm]
        leaves[nm] = torch.tensor(float(val), dtype=torch.float64, requires_grad=True)
        for term in th:
            if nm in th[term]:
                th[term][nm] = leaves[nm]
    params = orc.init_all(c["model"], th)
    center = torch.tensor(c["center"][frame], requires_grad=True)
    quat = torch.tensor(c["quat"][frame], requires_grad=True)
    t = orc.energy_terms(c["model"], center, quat, c["seq"], c["bonded"], c["pairs"], params, box=20.0, is_end=c["is_end"],
                         nt_type=c["nt_type"], stack_nt_type=stack_nt_type(c))
    (t * torch.as_tensor(cot)).sum().backward()
    return center.grad.numpy(), quat.grad.numpy(), {k: (0.0 if v.grad is None else float(v.grad)) for k, v in leaves.items()}


@pytest.mark.parametrize("name", GRAD_CASES)
def test_forces_and_theta_gradients_match_oracle(name):
    c = load_case(name)
    names = THETA[c["model"]]
    frame = 11
    rng = np.random.default_rng(3)
    cot = rng.uniform(0.5, 1.5, size=8)
    efn = energy_fn_of(c)
    theta = {}
    for nm in names:
        val = next(getattr(fn.params, nm) for fn in efn.energy_fns if nm in fn.params and getattr(fn.params, nm) is not None)
        theta[nm] = torch.tensor(float(val), dtype=torch.float64, requires_grad=True)
    efn_t = efn.with_params(theta)
    body = body_of(c, frame, requires_grad=True)
    terms = efn_t.compute_terms(body)
    w = torch.tensor(cot[: terms.shape[0]], device=DEV)
    (terms * w).sum().backward()
    want_c, want_q, want_th = _oracle_theta_grads(c, names, frame, cot)
    for got, want in ((body.center.grad.cpu().numpy(), want_c), (body.orientation.vec.grad.cpu().numpy(), want_q)):
        np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-7 * np.abs(want).max())
    for nm in names:
        got = 0.0 if theta[nm].grad is None else float(theta[nm].grad)  # unused under ss weights -> no graph
        assert np.isclose(got, want_th[nm], rtol=1e-6, atol=1e-8), (nm, got, want_th[nm])


@pytest.mark.parametrize("name", ["dna1_simple_helix", "dna2_half_charged", "rna2_helix_12bp", "na1_helix_rna_dna"])
def test_float32_within_1e4_of_oracle(name):
    c = load_case(name)
    efn = energy_fn_of(c)
    frame = 5
    body = body_of(c, frame, dtype=torch.float32, requires_grad=True)
    terms = efn.compute_terms(body)
    terms.sum().backward()
    want_c, want_q, _ = _oracle_theta_grads(c, [], frame, np.ones(8))
    params = orc.init_all(c["model"], theta_for(c))
    want_t = orc.energy_terms(c["model"], c["center"][frame], c["quat"][frame], c["seq"], c["bonded"], c["pairs"], params,
                              box=20.0, is_end=c["is_end"], nt_type=c["nt_type"], stack_nt_type=stack_nt_type(c)).numpy()
    scale = np.abs(want_t).max()
    np.testing.assert_allclose(terms.detach().cpu().numpy(), want_t[: terms.shape[0]], rtol=1e-4, atol=1e-4 * scale)
    np.testing.assert_allclose(body.center.grad.cpu().numpy(), want_c, rtol=1e-4, atol=1e-4 * np.abs(want_c).max())
    np.testing.assert_allclose(body.orientation.vec.grad.cpu().numpy(), want_q, rtol=1e-4, atol=1e-4 * np.abs(want_q).max())


def test_edge_cases_empty_and_padded_lists():
    c = load_case("dna1_simple_helix")
    efn = energy_fn_of(c)
    n = c["center"].shape[1]
    body = body_of(c, 0)
    full = efn.compute_terms(body).cpu().numpy()
    # padding entries (index N) interleaved and appended: identical energies
    pairs = np.asarray(c["pairs"])
    pad = np.full((2, 37), n, dtype=pairs.dtype)
    padded = np.concatenate([pad[:, :5], pairs[:, :40], pad[:, 5:20], pairs[:, 40:], pad[:, 20:]], axis=1)
    got = efn.with_props(unbonded_neighbors=padded).compute_terms(body).cpu().numpy()
    np.testing.assert_allclose(got, full, rtol=1e-13, atol=1e-13)
    # empty unbonded list: unbonded terms vanish, bonded terms unchanged
    empty = efn.with_props(unbonded_neighbors=np.zeros((2, 0), dtype=np.int32)).compute_terms(body).cpu().numpy()
    np.testing.assert_allclose(empty[:3], full[:3], rtol=1e-13)
    assert np.all(empty[3:] == 0.0)
    # CPU tensors are refused loudly (no fallback)
    from mythos_b200._lib import MythosB200Error

    with pytest.raises(MythosB200Error):
        efn(RigidBody(torch.tensor(c["center"][0]), Quaternion(torch.tensor(c["quat"][0]))))


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("name", ["dna1_simple_helix", "dna2_half_charged", "dna2_simple_coax", "rna2_helix_12bp",
                                  "na1_helix_dna_rna", "na1_helix_rna_dna", "na1_coax_rna"])
def test_list_kernels_equal_generic_kernels(name, dtype):
    """The phase-queued list kernels (large-system path: Debye / filter pass + short-range queue pass) against the
    one-thread-per-pair kernels on the golden frames: per-term energies, forces, dE/dquat and dE/dparams of all banks."""
    from mythos_b200 import _lib
    from mythos_b200.energy import functional
    from mythos_b200.energy import model as kmodel

    c = load_case(name)
    efn = energy_fn_of(c)
    plan = kmodel.plan_for(efn.energy_fns)
    cd = torch.tensor(c["center"][:20], dtype=dtype, device=DEV)
    qd = torch.tensor(c["quat"][:20], dtype=dtype, device=DEV)
    topo = plan.topology(cd.shape[1], cd.device)
    params = plan.device_params(cd.device, dtype)
    pairs = plan.pairs(cd.device, topo)
    cot = torch.tensor(np.random.default_rng(2).uniform(0.5, 1.5, size=(20, 8)), device=DEV, dtype=dtype)
    outs = []
    for flags in (_lib.FLAG_LIST_KERNEL, _lib.FLAG_GENERIC_KERNEL):
        outs.append(functional.energy_and_gradients(plan.model, topo, cd, qd, params, pairs, cot=cot, want_pos_grad=True,
                                                    want_param_grad=True, flags=flags))
    tol = 1e-11 if dtype == torch.float64 else 3e-4
    for got, want in zip(outs[0], outs[1]):
        got, want = got.cpu().numpy(), want.cpu().numpy()
        np.testing.assert_allclose(got, want, rtol=tol, atol=tol * np.abs(want).max())


@pytest.mark.parametrize("model", ["dna2", "na1"])
def test_tagged_lists_through_list_kernels_equal_plain_lists(model):
    """Large-system route of the AllPairs sentinel with forces (N >= 4096): two support-tagged neighbour builds (centres at
    the short-range cutoff, backbone sites at the Debye cutoff) feeding the list kernels, against one plain build at the
    interaction range through the list kernels and through the one-thread-per-pair kernels."""
    from mythos_b200 import _lib
    from mythos_b200.energy import dna2, functional, na1
    from mythos_b200.energy import model as kmodel
    from mythos_b200.input.topology import AllPairs
    from mythos_b200.utils import synthetic

    pattern = ((1, 1), (2, 2), (1, 2)) if model == "na1" else None
    s = synthetic.assembly(36, seed=4, nt_pattern=pattern)
    n = s.center.shape[0]
    assert n >= 4096
    efn = (na1 if model == "na1" else dna2).create_default_energy_fn(s.topology).with_props(unbonded_neighbors=AllPairs(n))
    plan = kmodel.plan_for(efn.energy_fns)
    cd = torch.tensor(s.center[None], device=DEV)
    qd = torch.tensor(s.quat[None], device=DEV)
    topo = plan.topology(n, cd.device)
    params = plan.device_params(cd.device, torch.float64)
    cot = torch.tensor(np.random.default_rng(5).uniform(0.5, 1.5, size=(1, 8)), device=DEV)
    outs = []
    for tagged, flags in ((True, 0), (False, 0), (False, _lib.FLAG_GENERIC_KERNEL)):
        src = plan.pairs(cd.device, topo)
        assert src.tag is not None
        src.tag_for_list_kernels = tagged
        outs.append(functional.energy_and_gradients(plan.model, topo, cd, qd, params, src, cot=cot, want_pos_grad=True,
                                                    want_param_grad=True, flags=flags))
        if tagged:
            assert src.tagged_capacity > 0 and src.last_split is not None
    for k in (0, 1):
        for got, want in zip(outs[k], outs[2]):
            got, want = got.cpu().numpy(), want.cpu().numpy()
            np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-10 * np.abs(want).max())
