"""Differentiable-through-time MD: the integrator's adjoint kernel against torch autograd through the
oracle step (exact to round-off: both are analytic), and the full trajectory gradient -- theta and initial state -- against
autograd through oracle step + oracle energy on the reference's 16-nt dna1 helix (the force's Hessian part comes from a
central difference of analytic forces: 1e-6)."""

import ctypes as C

import numpy as np
import pytest
import torch

from mythos_b200 import _lib
from mythos_b200.energy import dna1
from mythos_b200.rigid_body import Quaternion, RigidBody
from mythos_b200.simulators import adjoint
from oracle import langevin_oracle as lo
from oracle import oxdna_oracle as orc
from tests.golden_cases import load_case, theta_for
from tests.product_cases import topology_of

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
DT, KT = 5e-3, 296.15 * 0.1 / 300.0
GC, GQ = KT / 2.5, KT / 7.5


@pytest.mark.parametrize("phase", [0, 2])
def test_integrator_adjoint_matches_autograd_through_the_oracle_step(phase):
    rng = np.random.default_rng(3)
    n = 37
    c, pc, gc = (torch.tensor(rng.normal(size=(n, 3))) for _ in range(3))
    q = torch.tensor(rng.normal(size=(n, 4)))
    q = q / q.norm(dim=1, keepdim=True)
    pq, gq = (torch.tensor(rng.normal(size=(n, 4))) for _ in range(2))
    noise = torch.tensor(rng.normal(size=(n, 6)))
    lam_out = [torch.tensor(rng.normal(size=s)) for s in ((n, 3), (n, 4), (n, 3), (n, 4))]
    inertia = [1.0, 1.3, 0.8]

    leaves = [t.clone().requires_grad_(True) for t in (c, q, pc, pq, gc, gq)]
    out = lo.step_torch(*leaves, noise, DT, KT, GC, GQ, 1.2, inertia, kick=DT if phase == 2 else None)
    want = torch.autograd.grad(sum((o * l).sum() for o, l in zip(out, lam_out)), leaves)

    st = adjoint._Stepper(DT, KT, RigidBody(torch.tensor(GC, dtype=torch.float64), torch.tensor([GQ] * 3, dtype=torch.float64)), RigidBody(torch.tensor(1.2, dtype=torch.float64), torch.tensor(inertia, dtype=torch.float64)), torch.float64)
    dev = lambda t: t.to(DEV).contiguous()  # noqa: E731
    lam = [dev(t) for t in lam_out]
    lam_force = [torch.empty((n, 3), device=DEV, dtype=torch.float64), torch.empty((n, 4), device=DEV, dtype=torch.float64)]
    st.adjoint(dev(c), dev(q), dev(pc), dev(pq), dev(gc), dev(gq), dev(noise), phase, lam, lam_force)
    for got, w in zip(lam + lam_force, want):
        np.testing.assert_allclose(got.cpu().numpy(), w.numpy(), rtol=1e-10, atol=1e-12)
    # and the forward kernel agrees with the same oracle step
    cc, qq, ppc, ppq = dev(c), dev(q), dev(pc), dev(pq)
    st.forward(cc, qq, ppc, ppq, dev(gc), dev(gq), dev(noise), phase)
    for got, w in zip((cc, qq, ppc, ppq), out):
        np.testing.assert_allclose(got.cpu().numpy(), w.detach().numpy(), rtol=1e-11, atol=1e-12)


def test_trajectory_gradient_matches_autograd_through_oracle_md():
    case = load_case("dna1_simple_helix")
    top = topology_of(case)
    efn = dna1.create_default_energy_fn(top)
    n, steps = case["center"].shape[1], 4
    c0, q0 = torch.tensor(case["center"][0]), torch.tensor(case["quat"][0])
    rng = np.random.default_rng(0)
    noise = torch.tensor(rng.normal(size=(steps, n, 6)))
    names = ["eps_backbone", "a_stack", "eps_hb", "k_cross", "sigma_backbone"]
    base = efn.params_dict(include_dependent=False)
    theta = {k: torch.tensor(float(base[k]), dtype=torch.float64) for k in names}
    target = torch.tensor(rng.normal(size=(n, 3)))

    def loss_fn(traj):  # end-to-end distance-like loss on every stored frame, plus an orientation term
        return ((traj.center - target.to(traj.center.device)) ** 2).mean() + 0.3 * (traj.orientation.vec[-1, :, 0] ** 2).sum()

    loss, grads, traj, init_grad = adjoint.simulate_and_grad(
        efn, theta, RigidBody(c0.to(DEV), Quaternion(q0.to(DEV))), steps, loss_fn, dt=DT, kT=KT,
        gamma=RigidBody(torch.tensor(GC, dtype=torch.float64), torch.tensor([GQ] * 3, dtype=torch.float64)), noise=noise)

    # oracle: autograd through oracle step + oracle energy
    th = orc.default_theta("dna1")
    leaves = {}
    for nm in names:
        for term in th:
            if nm in th[term]:
                leaves.setdefault(nm, torch.tensor(float(th[term][nm]), dtype=torch.float64, requires_grad=True))
                th[term][nm] = leaves[nm]
    params = orc.init_all("dna1", th)
    c = c0.clone().requires_grad_(True)
    q = q0.clone().requires_grad_(True)
    cc, qq = c, q
    pc, pq = torch.zeros(n, 3, dtype=torch.float64), torch.zeros(n, 4, dtype=torch.float64)
    frames_c, frames_q = [], []
    for k in range(steps):
        e = orc.energy_terms("dna1", cc, qq, case["seq"], case["bonded"], case["pairs"], params).sum()
        g_c, g_q = torch.autograd.grad(e, [cc, qq], create_graph=True)
        cc, qq, pc, pq = lo.step_torch(cc, qq, pc, pq, g_c, g_q, noise[k], DT, KT, GC, GQ, 1.0, [1.0, 1.0, 1.0], kick=None if k == 0 else DT)
        frames_c.append(cc)
        frames_q.append(qq)

    class _T:
        center = torch.stack(frames_c)
        orientation = Quaternion(torch.stack(frames_q))

    want_loss = loss_fn(_T)
    want = torch.autograd.grad(want_loss, [c, q, *leaves.values()])
    np.testing.assert_allclose(traj.center.cpu().numpy(), _T.center.detach().numpy(), rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(float(loss), float(want_loss), rtol=1e-10)
    np.testing.assert_allclose(init_grad.center.cpu().numpy(), want[0].numpy(), rtol=1e-6, atol=1e-9 * float(want[0].abs().max()))
    np.testing.assert_allclose(init_grad.orientation.vec.cpu().numpy(), want[1].numpy(), rtol=1e-6, atol=1e-9 * float(want[1].abs().max()))
    for nm, w in zip(leaves, want[2:]):
        np.testing.assert_allclose(float(grads[nm]), float(w), rtol=1e-5, atol=1e-12, err_msg=nm)
