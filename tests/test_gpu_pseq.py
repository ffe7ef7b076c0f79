"""Probabilistic sequences on the GPU: one-hot distributions reproduce the discrete-sequence
energies (the reference's own integration check, e.g. dna2/tests/test_integration.py:150-157), and the energy of a random
distribution equals the brute-force expectation over every discrete sequence
(dna1/tests/test_expected_energies.py:162-328 does exactly this on a 4-bp helix), values and gradients."""

import itertools

import numpy as np
import pytest
import torch

from mythos_b200.energy import dna1, dna2
from mythos_b200.input import sequence_constraints as jd_sc
from mythos_b200.input.topology import from_strands
from mythos_b200.rigid_body import Quaternion, RigidBody
from mythos_b200.utils import synthetic
from tests.golden_cases import load_case
from tests.product_cases import energy_fn_of

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("name", ["dna1_simple_helix", "dna2_simple_helix", "dna1_seq_dep"])
def test_one_hot_pseq_equals_discrete_sequence(name):
    c = load_case(name)
    efn = energy_fn_of(c)
    n = c["center"].shape[1]
    body = RigidBody(torch.tensor(c["center"][:8], device=DEV), Quaternion(torch.tensor(c["quat"][:8], device=DEV)))
    want = efn.compute_terms_frames(body)
    for bps in (np.zeros((0, 2), dtype=np.int64), np.array([[k, n - 1 - k] for k in range(n // 2)])):
        sc = jd_sc.from_bps(n, bps)
        pseq = jd_sc.dseq_to_pseq(c["seq"], sc)
        got = efn.with_params(pseq=pseq, pseq_constraints=sc).compute_terms_frames(body)
        np.testing.assert_allclose(got.cpu().numpy(), want.cpu().numpy(), rtol=1e-12, atol=1e-12)
        e_map = efn.with_params(pseq=pseq, pseq_constraints=sc).map(body)
        np.testing.assert_allclose(e_map.cpu().numpy(), want.sum(1).cpu().numpy(), rtol=1e-12)


def _four_bp_system():
    c, q, _ = synthetic.ideal_duplex(4)
    rng = np.random.default_rng(4)
    top = from_strands(["ACGT", "ACGT"])
    frames = [synthetic.jitter(c, q, np.random.default_rng(100 + k)) for k in range(3)]
    cc = np.stack([f[0] for f in frames])
    qq = np.stack([f[1] for f in frames])
    return top, cc, qq, rng


@pytest.mark.parametrize("model", ["dna1", "dna2"])
def test_random_pseq_equals_brute_force_expectation_with_gradients(model):
    top, cc, qq, rng = _four_bp_system()
    mod = dna1 if model == "dna1" else dna2
    hb_w = torch.tensor(rng.random((4, 4)) + 0.1)
    st_w = torch.tensor(rng.random((4, 4)) + 0.5)
    efn = mod.create_default_energy_fn(top).with_params({"ss_hb_weights": hb_w, "ss_stack_weights": st_w})
    body = RigidBody(torch.tensor(cc, device=DEV), Quaternion(torch.tensor(qq, device=DEV)))
    sc = jd_sc.from_bps(8, np.array([[0, 7], [1, 6], [2, 5]]))
    up = rng.random((sc.n_unpaired, 4))
    up /= up.sum(1, keepdims=True)
    bp = rng.random((sc.n_bp, 4))
    bp /= bp.sum(1, keepdims=True)
    up_t = torch.tensor(up, requires_grad=True)
    bp_t = torch.tensor(bp, requires_grad=True)

    e = efn.with_params(pseq=(up_t, bp_t), pseq_constraints=sc).map(body)  # (3,)
    cot = torch.tensor([1.0, -0.5, 2.0], device=DEV, dtype=torch.float64)
    (e * cot).sum().backward()

    # brute force: every assignment of the 2 unpaired nucleotides and the 3 base-pair types
    slots = sc.n_unpaired + sc.n_bp
    energies, combos = [], []
    for idx in itertools.product(range(4), repeat=slots):
        dseq = np.zeros(8, dtype=np.int64)
        for k, nt in enumerate(sc.unpaired):
            dseq[nt] = idx[k]
        for k, (a, b) in enumerate(sc.bps):
            dseq[a], dseq[b] = jd_sc.BP_IDXS[idx[sc.n_unpaired + k]]
        with torch.no_grad():
            energies.append(efn.with_props(seq=torch.tensor(dseq)).map(body).cpu())
        combos.append(idx)
    E = torch.stack(energies)  # (1024, 3)
    up_b = torch.tensor(up, requires_grad=True)
    bp_b = torch.tensor(bp, requires_grad=True)
    idx = torch.tensor(combos)
    prob = torch.ones(len(combos), dtype=torch.float64)
    for k in range(sc.n_unpaired):
        prob = prob * up_b[k][idx[:, k]]
    for k in range(sc.n_bp):
        prob = prob * bp_b[k][idx[:, sc.n_unpaired + k]]
    expected = prob @ E
    np.testing.assert_allclose(e.detach().cpu().numpy(), expected.detach().numpy(), rtol=1e-10)
    (expected * cot.cpu()).sum().backward()
    # The two expressions are different extensions of the same function off the probability simplex (the product form
    # multiplies every pair energy by ALL slots' probabilities, the per-pair weights only by the slots the pair touches), so
    # their gradients differ by a constant per row and agree on the simplex's tangent space: compare with row means removed.
    tangent = lambda g: (g - g.mean(1, keepdim=True)).numpy()  # noqa: E731
    np.testing.assert_allclose(tangent(up_t.grad), tangent(up_b.grad), rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(tangent(bp_t.grad), tangent(bp_b.grad), rtol=1e-8, atol=1e-9)


def test_table_and_theta_gradients_with_pseq_match_finite_differences():
    top, cc, qq, rng = _four_bp_system()
    body = RigidBody(torch.tensor(cc, device=DEV), Quaternion(torch.tensor(qq, device=DEV)))
    sc = jd_sc.from_bps(8, np.array([[0, 7], [1, 6], [2, 5]]))
    up = torch.tensor(rng.dirichlet(np.ones(4), sc.n_unpaired))
    bp = torch.tensor(rng.dirichlet(np.ones(4), sc.n_bp))
    base = dna1.create_default_energy_fn(top)
    hb_w0 = torch.tensor(rng.random((4, 4)) + 0.1)

    def total(hb_w, eps_scale):
        efn = base.with_params({"ss_hb_weights": hb_w, "a_hb": 8.0 * eps_scale}).with_params(pseq=(up, bp), pseq_constraints=sc)
        return efn.map(body).sum()

    hb_w = hb_w0.clone().requires_grad_(True)
    scale = torch.tensor(1.0, dtype=torch.float64, requires_grad=True)
    total(hb_w, scale).backward()
    h = 1e-6
    for (a, b) in ((0, 3), (2, 1), (1, 1)):
        d = torch.zeros(4, 4, dtype=torch.float64)
        d[a, b] = h
        with torch.no_grad():
            fd = (float(total(hb_w0 + d, torch.tensor(1.0, dtype=torch.float64))) - float(total(hb_w0 - d, torch.tensor(1.0, dtype=torch.float64)))) / (2 * h)
        np.testing.assert_allclose(float(hb_w.grad[a, b]), fd, rtol=1e-6, atol=1e-8)
    with torch.no_grad():
        fd = (float(total(hb_w0, torch.tensor(1.0 + h, dtype=torch.float64))) - float(total(hb_w0, torch.tensor(1.0 - h, dtype=torch.float64)))) / (2 * h)
    np.testing.assert_allclose(float(scale.grad), fd, rtol=1e-6, atol=1e-8)
