"""DiffTRe pass on the GPU: all-pairs semantics via per-frame cell lists, fused E + dE/dparams, reweighting.

Compares with the oracle on a small batch of frames of the benchmark system (N = 2040), and checks the
size-independent properties the pass must have at full size (additivity over frames, weights sum to one,
gradient of a frame-linear loss == g-weighted sum of per-frame gradients).
"""

import numpy as np
import pytest
import torch

from mythos_b200.energy import dna2
from mythos_b200.optimization import objective
from mythos_b200.rigid_body import Quaternion, RigidBody
from mythos_b200.simulators.io import SimulatorTrajectory
from mythos_b200.utils import synthetic
from oracle import oxdna_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def workload():
    s = synthetic.assembly(17, seed=1)
    c, q = synthetic.rejittered_frames(s, 6)
    return s, c, q


def test_map_matches_oracle_and_theta_gradient(workload):
    s, c, q = workload
    top = s.topology
    efn = dna2.create_default_energy_fn(top)
    names = ["eps_backbone", "a_stack", "eps_hb", "k_cross", "q_eff", "lambda_factor", "sigma_backbone", "kt"]
    base = efn.params_dict(include_dependent=False)
    theta = {k: torch.tensor(float(base[k]), dtype=torch.float64, requires_grad=True) for k in names}
    states = RigidBody(torch.tensor(c, device=DEV), Quaternion(torch.tensor(q, device=DEV)))
    e = efn.with_params(theta).map(states)
    g = torch.tensor(np.random.default_rng(0).standard_normal(len(c)), device=DEV)
    (e * g).sum().backward()

    th = orc.default_theta("dna2")
    leaves = {}
    for nm in names:
        for term in th:
            if nm in th[term]:
                leaves.setdefault(nm, torch.tensor(float(th[term][nm]), dtype=torch.float64, requires_grad=True))
                th[term][nm] = leaves[nm]
    params = orc.init_all("dna2", th)
    want = []
    for f in range(len(c)):
        pairs = orc.neighbor_pairs(c[f], top.bonded_neighbors, 3.3, 0.0)
        want.append(orc.energy_terms("dna2", c[f], q[f], top.seq, top.bonded_neighbors, pairs, params, is_end=top.is_end).sum())
    want = torch.stack(want)
    (want * g.cpu()).sum().backward()
    np.testing.assert_allclose(e.detach().cpu().numpy(), want.detach().numpy(), rtol=1e-9)
    for nm in names:
        assert np.isclose(float(theta[nm].grad), float(leaves[nm].grad), rtol=1e-6, atol=1e-9), nm


def test_weights_and_neff_kats():
    # mythos/optimization/tests/test_objective.py:187-205: equal energies -> uniform weights, n_eff = 1
    e = torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64, device=DEV)
    w, neff = objective.compute_weights_and_neff(1.0, e, e)
    np.testing.assert_allclose(w.cpu().numpy(), [1 / 3] * 3)
    np.testing.assert_allclose(float(neff), 1.0)
    # general case against the oracle's formula, including the gradient through weights and n_eff
    rng = np.random.default_rng(1)
    en = torch.tensor(rng.normal(size=257), dtype=torch.float64, device=DEV, requires_grad=True)
    er = torch.tensor(rng.normal(size=257), dtype=torch.float64, device=DEV)
    beta = torch.tensor(rng.uniform(5, 12, size=257), dtype=torch.float64, device=DEV)
    w, neff = objective.compute_weights_and_neff(beta, en, er)
    obs = torch.tensor(rng.normal(size=257), device=DEV)
    ((w * obs).sum() + 0.3 * neff).backward()
    en2 = en.detach().cpu().clone().requires_grad_(True)
    w2, neff2 = orc.weights_and_neff(beta.cpu(), en2, er.cpu())
    ((w2 * obs.cpu()).sum() + 0.3 * neff2).backward()
    np.testing.assert_allclose(w.detach().cpu().numpy(), w2.detach().numpy(), rtol=1e-12)
    np.testing.assert_allclose(float(neff.detach()), float(neff2.detach()), rtol=1e-12)
    np.testing.assert_allclose(en.grad.cpu().numpy(), en2.grad.numpy(), rtol=1e-9, atol=1e-14)
    # multi-temperature segments (test_objective.py:483-497)
    temp = torch.tensor([0.1] * 100 + [0.11] * 157, dtype=torch.float64, device=DEV)
    m = objective.compute_min_segment_neff(temp, en.detach(), er)
    a = float(orc.weights_and_neff(torch.tensor(10.0), en2[:100].detach(), er.cpu()[:100])[1])
    b = float(orc.weights_and_neff(torch.tensor(1 / 0.11), en2[100:].detach(), er.cpu()[100:])[1])
    np.testing.assert_allclose(m, min(a, b), rtol=1e-7)


def test_difftre_objective_state_machine_and_gradient(workload):
    s, c, q = workload
    efn = dna2.create_default_energy_fn(s.topology)
    kT = float(dna2.default_configs()[0]["kT"])
    F = len(c)
    traj = SimulatorTrajectory(
        center=torch.tensor(c, device=DEV), orientation=Quaternion(torch.tensor(q, device=DEV)),
        temperature=torch.full((F,), kT, dtype=torch.float64, device=DEV))
    obs = torch.tensor(np.random.default_rng(2).normal(size=F), device=DEV)

    def loss_fn(ref_states, weights, energy_fn, opt_params, observables):
        m = (weights * obs).sum()
        return (m - 0.1) ** 2, (("obs", m), None)

    objv = objective.DiffTReObjective(name="t", required_observables=("traj",), grad_or_loss_fn=loss_fn, energy_fn=efn,
                                      min_n_eff_factor=0.5)
    out = objv.calculate({}, {})
    assert not out.is_ready and out.needs_update == ("traj",)
    theta = {"eps_hb": torch.tensor(1.0678, dtype=torch.float64), "a_stack": torch.tensor(6.0, dtype=torch.float64)}
    out = objv.calculate({"traj": traj}, theta)
    assert out.is_ready and out.state["opt_steps"] == 1 and abs(out.observables["neff"] - 1.0) < 1e-12
    # at theta == theta_ref the weights are uniform: dL/dtheta = 2 (m - 0.1) * d<O>_w/dtheta with
    # d<O>_w/dtheta = -beta * Cov_uniform(O, dE/dtheta)
    leaves = {k: v.clone().requires_grad_(True) for k, v in theta.items()}
    e = efn.with_params(leaves).map(traj)
    m = float(obs.mean())
    for k in theta:
        (ge,) = torch.autograd.grad(e, leaves[k], grad_outputs=torch.eye(F, device=DEV, dtype=torch.float64)[0] * 0 + 1, retain_graph=True)
        per = []
        for f in range(F):
            (gf,) = torch.autograd.grad(e[f], leaves[k], retain_graph=True)
            per.append(float(gf))
        per = np.array(per)
        cov = float((obs.cpu().numpy() * per).mean() - obs.cpu().numpy().mean() * per.mean())
        want = 2 * (m - 0.1) * (-(1 / kT) * cov)
        assert np.isclose(float(out.grads[k]), want, rtol=1e-6, atol=1e-12), (k, float(out.grads[k]), want)
    # a far-away reference makes n_eff collapse -> new trajectory requested
    far = {"eps_hb": torch.tensor(1.4, dtype=torch.float64), "a_stack": torch.tensor(6.0, dtype=torch.float64)}
    out2 = objective.DiffTReObjective(name="t", required_observables=("traj",), grad_or_loss_fn=loss_fn, energy_fn=efn,
                                      min_n_eff_factor=0.95).calculate({"traj": traj}, far, reference_opt_params=theta)
    assert not out2.is_ready and out2.state == {"opt_steps": 0}


def test_frame_additivity_at_scale(workload):
    """Size-independent property used at the benchmark size: the energies of a batch are those of its halves."""
    s, _, _ = workload
    c, q = synthetic.rejittered_frames(s, 700, seed0=5000)
    efn = dna2.create_default_energy_fn(s.topology)
    cd, qd = torch.tensor(c, device=DEV), torch.tensor(q, device=DEV)
    e_all = efn.map(RigidBody(cd, Quaternion(qd)))
    e_a = efn.map(RigidBody(cd[:333], Quaternion(qd[:333])))
    e_b = efn.map(RigidBody(cd[333:], Quaternion(qd[333:])))
    np.testing.assert_allclose(e_all.cpu().numpy(), torch.cat([e_a, e_b]).cpu().numpy(), rtol=1e-11)
    assert torch.isfinite(e_all).all()


@pytest.mark.parametrize("model", ["dna1", "dna2"])
def test_screens_are_supersets_on_disordered_frames(model):
    """Strongly perturbed duplexes (positions +-0.12, rotations ~0.5 rad: base pairs half broken, excluded-volume overlaps,
    angles all over their windows): the frame kernel's screens (squared cutoffs, float32-rsqrt cosine windows, six-window
    hydrogen-bond queue, dense excluded-volume queue) may only drop pairs whose terms are exactly zero, so every route of the
    frame-resident kernel must reproduce the one-thread-per-pair kernel term by term and J row by J row."""
    from mythos_b200 import _lib
    from mythos_b200.energy import dna1, functional
    from mythos_b200.energy import model as kmodel

    s = synthetic.assembly(4, seed=11)
    rng = np.random.default_rng(5)
    frames = [synthetic.jitter(s.center, s.quat, rng, sigma_pos=0.12, sigma_rot=0.5) for _ in range(5)]
    c, q = np.stack([f[0] for f in frames]), np.stack([f[1] for f in frames])
    from mythos_b200.input.topology import AllPairs

    efn = (dna1 if model == "dna1" else dna2).create_default_energy_fn(s.topology).with_props(unbonded_neighbors=AllPairs(s.center.shape[0]))
    plan = kmodel.plan_for(efn.energy_fns)
    cd, qd = torch.tensor(c, device=DEV), torch.tensor(q, device=DEV)
    topo = plan.topology(cd.shape[1], cd.device)
    params = plan.device_params(cd.device, torch.float64)
    cot = torch.tensor(np.random.default_rng(2).uniform(0.5, 1.5, size=(len(c), 8)), device=DEV)
    outs = {}
    for name, flags, in_kernel, tagged in (("pair kernel", _lib.FLAG_GENERIC_KERNEL, False, False), ("in-kernel cells", 0, True, False),
                                           ("plain lists", 0, False, False), ("tagged lists", 0, False, True)):
        src = plan.pairs(cd.device, topo)
        src.in_kernel = in_kernel
        if not tagged:
            src.tag = None
        terms, _, _, J = functional.energy_and_gradients(plan.model, topo, cd, qd, params, src, cot=cot, want_pos_grad=False,
                                                         want_param_grad=True, per_frame_param_grad=True, flags=flags)
        src.verify()
        outs[name] = (terms.cpu().numpy(), J.cpu().numpy())
    ref_t, ref_j = outs["pair kernel"]
    assert np.isfinite(ref_t).all() and np.abs(ref_t[:, 3]).max() > 0 and np.abs(ref_t[:, 4]).max() > 0  # excluded volume and hydrogen bonding are active
    for name in ("in-kernel cells", "plain lists", "tagged lists"):
        np.testing.assert_allclose(outs[name][0], ref_t, rtol=1e-10, atol=1e-10 * np.abs(ref_t).max(), err_msg=name)
        np.testing.assert_allclose(outs[name][1], ref_j, rtol=1e-9, atol=1e-10 * np.abs(ref_j).max(), err_msg=name)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_frame_resident_kernel_equals_pair_kernel(workload, dtype):
    """The frame-resident kernel (one CTA per frame, queues in shared memory) and the one-thread-per-pair kernel are
    two schedules of the same per-pair math: energies and dE/dparams rows must agree to rounding."""
    from mythos_b200 import _lib
    from mythos_b200.energy import functional
    from mythos_b200.energy import model as kmodel

    s, c, q = workload
    efn = dna2.create_default_energy_fn(s.topology)
    plan = kmodel.plan_for(efn.energy_fns)
    cd, qd = torch.tensor(c, device=DEV, dtype=dtype), torch.tensor(q, device=DEV, dtype=dtype)
    topo = plan.topology(cd.shape[1], cd.device)
    params = plan.device_params(cd.device, dtype)
    cot = torch.tensor(np.random.default_rng(1).uniform(0.5, 1.5, size=(len(c), 8)), device=DEV, dtype=dtype)
    outs = []
    for flags, in_kernel, tagged in ((0, True, False), (_lib.FLAG_GENERIC_KERNEL, False, False), (0, False, False), (0, False, True)):
        # in_kernel: frame kernel finds its own pairs (shared-memory cell list); GENERIC: device lists + pair kernels;
        # plain: device lists streamed through the frame kernel; tagged (the default): the neighbour build keeps only pairs
        # inside the support of some term and tags which, the frame kernel queues them without touching coordinates
        src = plan.pairs(cd.device, topo)
        src.in_kernel = in_kernel
        assert src.tag is not None
        if not tagged:
            src.tag = None
        terms, _, _, J = functional.energy_and_gradients(plan.model, topo, cd, qd, params, src, cot=cot, want_pos_grad=False,
                                                         want_param_grad=True, per_frame_param_grad=True, flags=flags)
        if tagged:
            assert src.slot_geometry is not None and src.tag is not None  # the tagged (warp-slot) route really ran
        outs.append((terms.cpu().numpy(), J.cpu().numpy()))
    tol = 1e-11 if dtype == torch.float64 else 2e-4
    np.testing.assert_allclose(outs[0][0], outs[1][0], rtol=tol, atol=tol * np.abs(outs[1][0]).max())
    np.testing.assert_allclose(outs[0][1], outs[1][1], rtol=tol, atol=tol * np.abs(outs[1][1]).max())
    for k in (2, 3):
        np.testing.assert_allclose(outs[k][0], outs[1][0], rtol=tol, atol=tol * np.abs(outs[1][0]).max())
        np.testing.assert_allclose(outs[k][1], outs[1][1], rtol=tol, atol=tol * np.abs(outs[1][1]).max())


@pytest.mark.parametrize("in_kernel", [False, True])
@pytest.mark.parametrize("box", [None, 9.0, 30.0])
def test_all_pairs_sentinel_equals_explicit_all_pairs_list(box, in_kernel):
    """AllPairs (device cell lists, or the in-kernel cell list, at the interaction range) == the reference's explicit
    N(N-1)/2 list, in free space and in periodic boxes (a 9.0 box is smaller than the stencil along each axis:
    single-cell axes)."""
    from mythos_b200 import space
    from mythos_b200.input.topology import AllPairs

    s = synthetic.assembly(2, seed=6)
    top = s.topology
    disp = space.periodic(box)[0] if box else space.free()[0]
    efn = dna2.create_default_energy_fn(top, displacement_fn=disp)  # N = 240 < 512: explicit list
    c, q = synthetic.rejittered_frames(s, 5, seed0=77)
    states = RigidBody(torch.tensor(c, device=DEV), Quaternion(torch.tensor(q, device=DEV)))
    want = efn.compute_terms_frames(states)
    got = efn.with_props(unbonded_neighbors=AllPairs(top.n_nucleotides, in_kernel=in_kernel)).compute_terms_frames(states)
    np.testing.assert_allclose(got.cpu().numpy(), want.cpu().numpy(), rtol=1e-10, atol=1e-10)


def test_pinned_host_frames_are_streamed_and_give_identical_results(workload, no_pair_list_cache):
    """``map`` over frames that live in pinned host memory (streamed to the device chunk by chunk on a copy stream) must
    reproduce the device-resident result (1e-13), energies and theta-gradients; pageable host memory is refused."""
    from mythos_b200 import _lib
    from mythos_b200.energy import functional

    s, c, q = workload
    efn = dna2.create_default_energy_fn(s.topology)
    theta = {"eps_hb": torch.tensor(float(efn.params_dict(include_dependent=False)["eps_hb"]), dtype=torch.float64, requires_grad=True)}
    old = functional.FRAME_CHUNK
    functional.FRAME_CHUNK = 2  # three chunks for the six test frames
    try:
        outs = []
        for host in (False, True):
            th = {k: v.detach().clone().requires_grad_(True) for k, v in theta.items()}
            cc, qq = torch.tensor(c), torch.tensor(q)
            cc, qq = (cc.pin_memory(), qq.pin_memory()) if host else (cc.to(DEV), qq.to(DEV))
            e = efn.with_params(th).map(RigidBody(cc, Quaternion(qq)))
            assert e.is_cuda
            (e * torch.arange(1, e.shape[0] + 1, device=e.device, dtype=e.dtype)).sum().backward()
            outs.append((e.detach().cpu().numpy(), float(th["eps_hb"].grad)))
        # (streamed passes cut their chunks differently -- ramping up from one wave -- so the per-chunk list layout and with
        # it the summation order inside a frame may differ in the last bit)
        np.testing.assert_allclose(outs[0][0], outs[1][0], rtol=1e-13)
        np.testing.assert_allclose(outs[0][1], outs[1][1], rtol=1e-12)
        with pytest.raises(_lib.MythosB200Error):
            efn.map(RigidBody(torch.tensor(c), Quaternion(torch.tensor(q))))
    finally:
        functional.FRAME_CHUNK = old


def test_loss_and_grad_with_prefetched_pinned_frames_equals_device_frames(workload, no_pair_list_cache):
    """``compute_loss_and_grad`` starts the copy of the first chunk of pinned host frames before the theta chain
    (``functional.prefetch_frames``); loss and gradients must be those of device-resident frames, also when the pass
    that follows uses other buffers than the prefetched ones (the stale copy is dropped)."""
    from mythos_b200.energy import functional
    from mythos_b200.optimization import objective

    s, c, q = workload
    efn = dna2.create_default_energy_fn(s.topology)
    base = efn.params_dict(include_dependent=False)
    theta = {k: torch.tensor(float(base[k]), dtype=torch.float64) for k in ("eps_hb", "a_stack", "q_eff")}
    F = c.shape[0]
    beta = torch.full((F,), 10.0, dtype=torch.float64, device=DEV)
    obs = torch.linspace(-1.0, 1.0, F, dtype=torch.float64, device=DEV)

    def loss_fn(ref_states, weights, energy_fn, opt_params, observables):
        m = (weights * obs).sum()
        return m, (("obs", m), None)

    old = functional.FRAME_CHUNK
    functional.FRAME_CHUNK = 4  # two chunks for the six test frames
    try:
        res = []
        for host in (False, True):
            cc, qq = torch.tensor(c), torch.tensor(q)
            cc, qq = (cc.pin_memory(), qq.pin_memory()) if host else (cc.to(DEV), qq.to(DEV))
            states = RigidBody(cc, Quaternion(qq))
            with torch.no_grad():
                e_ref = efn.map(states) + obs * 0.01
            (l, _), g = objective.compute_loss_and_grad(theta, efn, beta, loss_fn, states, e_ref, [])
            res.append((float(l), {k: float(v) for k, v in g.items()}))
            if host:
                assert not functional._PREFETCHED  # consumed by the pass
                functional.prefetch_frames(cc.clone().pin_memory(), qq)  # a copy nobody picks up ...
                assert functional._PREFETCHED
                e_again = efn.map(states)  # ... is dropped by the next streamed pass over other buffers
                assert not functional._PREFETCHED
                np.testing.assert_allclose(e_again.cpu().numpy(), (e_ref - obs * 0.01).cpu().numpy(), rtol=1e-13)
        # beta * E ~ 1e4 amplifies the last-bit summation-order noise of the energies in the weights: 1e-9, not bit-equal
        np.testing.assert_allclose(res[0][0], res[1][0], rtol=1e-9)
        for k in res[0][1]:
            np.testing.assert_allclose(res[0][1][k], res[1][1][k], rtol=1e-8, atol=1e-12)
    finally:
        functional.FRAME_CHUNK = old


def test_deferred_verification_reports_an_overflowed_pass_and_the_repeat_is_right(workload, no_pair_list_cache):
    """``functional.deferred_verification``: the evaluation does not read the overflow flags itself; the caller does after
    its own launches.  Slots that are too narrow must be reported (``ok()`` False, geometry re-fitted), and the repeated
    pass must give the energies and dE/dparams rows of an ordinary pass."""
    from mythos_b200 import _lib
    from mythos_b200.energy import functional
    from mythos_b200.energy import model as kmodel

    s, c, q = workload
    efn = dna2.create_default_energy_fn(s.topology)
    plan = kmodel.plan_for(efn.energy_fns)
    cd, qd = torch.tensor(c, device=DEV), torch.tensor(q, device=DEV)
    topo = plan.topology(cd.shape[1], DEV)
    params = plan.device_params(DEV, torch.float64)
    ones = torch.ones((cd.shape[0], _lib.N_TERMS), dtype=torch.float64, device=DEV)

    def run(src):
        t, _, _, J = functional.energy_and_gradients(plan.model, topo, cd, qd, params, src, cot=ones, want_pos_grad=False,
                                                     want_param_grad=True, per_frame_param_grad=True)
        return t, J

    t_ref, j_ref = run(plan.pairs(DEV, topo))
    src = plan.pairs(DEV, topo)
    try:
        src.slot_geometry = ((8, 32), (8, 32))  # far too narrow
        with functional.deferred_verification() as checks:
            run(src)
            assert src._pending  # nothing was read yet
        assert not checks.ok()
        assert src.slot_geometry != ((8, 32), (8, 32))
        for _ in range(8):  # statistics of an overflowed pass can be truncated: the re-fit may take more than one repeat
            with functional.deferred_verification() as checks:
                t2, j2 = run(src)
            if checks.ok():
                break
        else:
            raise AssertionError("slot geometry did not converge")
        np.testing.assert_allclose(t2.cpu().numpy(), t_ref.cpu().numpy(), rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(j2.cpu().numpy(), j_ref.cpu().numpy(), rtol=1e-10, atol=1e-10)
    finally:
        functional._SIZING.clear()


def test_loss_and_grad_recovers_when_fresh_sources_start_from_capacities_that_overflow(workload, no_pair_list_cache):
    """``compute_loss_and_grad`` rebuilds the energy function -- and with it the pair source -- on every repeat.  What a
    source learns when a list overflows (capacities, slot geometry, float64 fallback) must survive in the sizing memo, or
    the repeat overflows on the same frame forever (round-1 advisor finding).  Poison the memo with sizes that are far too
    small, and with a system too extended for the float32 tagged builds, and require the public call to converge to the
    result of a clean run."""
    from mythos_b200.energy import functional
    from mythos_b200.energy import model as kmodel
    from mythos_b200.input.topology import AllPairs

    s, c, q = workload
    n = s.center.shape[0]
    efn = dna2.create_default_energy_fn(s.topology).with_props(unbonded_neighbors=AllPairs(n))
    names = ["eps_hb", "k_cross", "q_eff"]
    base = efn.params_dict(include_dependent=False)
    theta = {k: torch.tensor(float(base[k]), dtype=torch.float64) for k in names}
    F = len(c)
    kT = 0.0987
    beta = torch.full((F,), 1.0 / kT, dtype=torch.float64, device=DEV)
    obs = torch.tensor(np.random.default_rng(3).standard_normal(F), device=DEV)

    def loss_fn(ref_states, weights, energy_fn, opt_params, observables):
        m = (weights * obs).sum()
        return m, (("obs", m), None)

    def run(cc):
        states = SimulatorTrajectory(center=torch.tensor(cc, device=DEV), orientation=Quaternion(torch.tensor(q, device=DEV)),
                                     temperature=torch.full((F,), kT, dtype=torch.float64, device=DEV))
        with torch.no_grad():
            e_ref = efn.map(states)
        return objective.compute_loss_and_grad(theta, efn, beta, loss_fn, states, e_ref + 0.01 * obs, [])

    try:
        functional._SIZING.clear()
        (l0, _), g0 = run(c)
        # (a) poisoned sizes: every fresh source starts from them
        probe = kmodel.plan_for(efn.energy_fns).pairs(torch.device(DEV), kmodel.plan_for(efn.energy_fns).topology(n, torch.device(DEV)))
        probe._load_memo(torch.device(DEV), n)
        functional._SIZING.update(probe._memo_key, slot_geometry=((8, 32), (8, 32)), capacity=64, tagged_capacity=64)
        (l1, _), g1 = run(c)
        assert np.isclose(float(l1), float(l0), rtol=1e-10)
        for k in names:
            assert np.isclose(float(g1[k]), float(g0[k]), rtol=1e-8, atol=1e-12), k
        assert functional._SIZING.get(probe._memo_key)["slot_geometry"] != ((8, 32), (8, 32))
        # (b) extent beyond 1500 length units: the float64 fallback must be remembered across the rebuilt sources
        functional._SIZING.clear()
        c2 = c.copy()
        c2[:, n // 2:, 0] += 4000.0
        (l2, _), g2 = run(c2)
        assert np.isfinite(float(l2)) and functional._SIZING.get(probe._memo_key).get("tag_float32") is False
    finally:
        functional._SIZING.clear()


def test_tagged_float32_builds_are_safe_far_from_the_origin_and_for_extended_systems(no_pair_list_cache):
    """The support-tagged builds run in float32 on recentred coordinates: a trajectory far from the origin must give the
    energies of the same trajectory at the origin, and a system too extended for float32 (two duplexes 4000 length units
    apart) must fall back to float64 builds and still agree with plain lists."""
    from mythos_b200.energy import functional
    from mythos_b200.energy import model as kmodel
    from mythos_b200.input.topology import AllPairs

    s = synthetic.assembly(2, seed=6)
    n = s.center.shape[0]
    efn = dna2.create_default_energy_fn(s.topology).with_props(unbonded_neighbors=AllPairs(n))
    c, q = synthetic.rejittered_frames(s, 4, seed0=21)
    here = efn.map(RigidBody(torch.tensor(c, device=DEV), Quaternion(torch.tensor(q, device=DEV)))).cpu().numpy()
    far = efn.map(RigidBody(torch.tensor(c + np.array([1.0e5, -3.0e4, 7.0e3]), device=DEV), Quaternion(torch.tensor(q, device=DEV)))).cpu().numpy()
    np.testing.assert_allclose(far, here, rtol=1e-8)

    c2 = c.copy()
    c2[:, n // 2:, 0] += 4000.0  # second duplex far away: extent > 1500 -> float64 builds
    plan = kmodel.plan_for(efn.energy_fns)
    cd, qd = torch.tensor(c2, device=DEV), torch.tensor(q, device=DEV)
    topo = plan.topology(n, cd.device)
    params = plan.device_params(cd.device, torch.float64)
    src = plan.pairs(cd.device, topo)
    got = functional.energy_and_gradients(plan.model, topo, cd, qd, params, src, want_pos_grad=False)[0].cpu().numpy()
    assert src.tag is not None and src.tag_float32 is False
    plain = plan.pairs(cd.device, topo)
    plain.tag = None
    want = functional.energy_and_gradients(plan.model, topo, cd, qd, params, plain, want_pos_grad=False)[0].cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-11, atol=1e-11)


@pytest.mark.parametrize("model", ["dna1", "rna2"])
def test_all_pairs_route_for_models_without_and_with_debye(model):
    """The default AllPairs route (support-tagged warp-slot lists + frame kernel) for oxDNA1 (no Debye-Hueckel: only the
    short-range build runs) and RNA2, against the explicit all-pairs list through the one-thread-per-pair kernels."""
    import mythos_b200.energy.dna1 as dna1
    import mythos_b200.energy.rna2 as rna2
    from mythos_b200.input.topology import AllPairs, unbonded_pairs

    s = synthetic.assembly(5, seed=9)
    top = s.topology
    n = top.n_nucleotides
    mod = dna1 if model == "dna1" else rna2
    efn = mod.create_default_energy_fn(top)
    c, q = synthetic.rejittered_frames(s, 3, seed0=31)
    states = RigidBody(torch.tensor(c, device=DEV), Quaternion(torch.tensor(q, device=DEV)))
    got = efn.with_props(unbonded_neighbors=AllPairs(n)).compute_terms_frames(states).cpu().numpy()
    explicit = np.ascontiguousarray(unbonded_pairs(n, top.bonded_neighbors).T)
    want = efn.with_props(unbonded_neighbors=explicit).compute_terms_frames(states).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-10 * np.abs(want).max())


def test_map_of_a_system_too_large_for_the_frame_kernel_falls_back_to_the_list_kernels():
    """N = 4320 does not fit the frame-resident kernel's shared memory in float64: the default AllPairs route must fall back
    (plain device lists + list kernels) and still give the explicit-route energies and dE/dtheta."""
    from mythos_b200 import _lib
    from mythos_b200.energy import functional
    from mythos_b200.energy import model as kmodel
    from mythos_b200.input.topology import AllPairs

    s = synthetic.assembly(36, seed=4)
    n = s.center.shape[0]
    efn = dna2.create_default_energy_fn(s.topology).with_props(unbonded_neighbors=AllPairs(n))
    c, q = synthetic.rejittered_frames(s, 2, seed0=3)
    th = {"k_cross": torch.tensor(float(efn.params_dict(include_dependent=False)["k_cross"]), dtype=torch.float64, requires_grad=True)}
    e = efn.with_params(th).map(RigidBody(torch.tensor(c, device=DEV), Quaternion(torch.tensor(q, device=DEV))))
    e.sum().backward()
    plan = kmodel.plan_for(efn.energy_fns)
    cd, qd = torch.tensor(c, device=DEV), torch.tensor(q, device=DEV)
    topo = plan.topology(n, cd.device)
    src = plan.pairs(cd.device, topo)
    src.tag = None
    terms, _, _, dp = functional.energy_and_gradients(plan.model, topo, cd, qd, plan.device_params(cd.device, torch.float64), src,
                                                      want_pos_grad=False, want_param_grad=True, flags=_lib.FLAG_GENERIC_KERNEL)
    np.testing.assert_allclose(e.detach().cpu().numpy(), terms.sum(1).cpu().numpy(), rtol=1e-11)
    k = _lib.param_names().index("cross_stacking.k_cross")
    assert np.isclose(float(th["k_cross"].grad), float(dp[k]), rtol=1e-9)


def test_loss_and_grad_through_the_replayed_theta_chain_equals_the_eager_chain(workload):
    """``compute_loss`` binds theta through the recorded tape (``theta_tape.bind``); with the replay switched off it runs
    the eager ``with_params`` chain.  Loss, n_eff and every gradient (all default optimisable parameters, plus a direct
    theta term in the loss) must agree; a loss function that calls the bound energy function gets the real object."""
    from mythos_b200.energy import theta_tape
    from mythos_b200.optimization import objective

    s, c, q = workload
    efn = dna2.create_default_energy_fn(s.topology)
    theta = {k: torch.as_tensor(v, dtype=torch.float64) * 1.01 for k, v in efn.opt_params().items()}
    F = c.shape[0]
    beta = torch.full((F,), 10.0, dtype=torch.float64, device=DEV)
    obs = torch.linspace(-1.0, 1.0, F, dtype=torch.float64, device=DEV)
    states = RigidBody(torch.tensor(c, device=DEV), Quaternion(torch.tensor(q, device=DEV)))
    seen = []

    def loss_fn(ref_states, weights, energy_fn, opt_params, observables):
        seen.append(float(energy_fn.params_dict()["eps_hb"]))  # answered by the materialised with_params result
        m = (weights * obs).sum() + 1e-3 * opt_params["eps_hb"] ** 2
        return m, (("obs", m), None)

    with torch.no_grad():
        e_ref = efn.map(states) + obs * 0.01
    res = []
    for enabled in (True, False):
        theta_tape.ENABLED = enabled
        try:
            (l, (neff, _, e_new)), g = objective.compute_loss_and_grad(theta, efn, beta, loss_fn, states, e_ref, [])
        finally:
            theta_tape.ENABLED = True
        res.append((float(l), float(neff), e_new.cpu().numpy(), {k: v.detach().cpu().numpy() for k, v in g.items()}))
    assert seen[0] == seen[1] == float(theta["eps_hb"])
    np.testing.assert_allclose(res[0][2], res[1][2], rtol=1e-13)
    np.testing.assert_allclose(res[0][0], res[1][0], rtol=1e-10)
    np.testing.assert_allclose(res[0][1], res[1][1], rtol=1e-10)
    scale = max(float(np.abs(v).max()) for v in res[1][3].values())
    for k in theta:
        np.testing.assert_allclose(res[0][3][k], res[1][3][k], rtol=1e-8, atol=1e-10 * scale, err_msg=k)


def test_remembered_pair_lists_give_the_same_pass_and_are_invalidated_correctly(workload):
    """``functional._PairListCache``: a second pass over the same frame tensors launches no neighbour kernels and returns
    the energies / gradients of a rebuilt pass (1e-12: the lists are supersets built with a margin, the kernels apply the
    exact supports); parameters that push a cutoff beyond the remembered one, an in-place update of the frames, and new
    tensor objects all trigger a rebuild; ``DiffTReObjective.calculate`` reaches the cache across calls."""
    from mythos_b200.energy import functional
    from mythos_b200.optimization import objective

    s, c, q = workload
    efn = dna2.create_default_energy_fn(s.topology)
    F = c.shape[0]
    cc, qq = torch.tensor(c, device=DEV), torch.tensor(q, device=DEV)
    states = RigidBody(cc, Quaternion(qq))
    theta = {k: torch.tensor(float(v), dtype=torch.float64) for k, v in efn.opt_params().items() if k in ("eps_hb", "a_stack", "q_eff", "lambda_factor")}
    beta = torch.full((F,), 10.0, dtype=torch.float64, device=DEV)
    obs = torch.linspace(-1.0, 1.0, F, dtype=torch.float64, device=DEV)

    def loss_fn(ref_states, weights, energy_fn, opt_params, observables):
        m = (weights * obs[-weights.shape[0]:]).sum()
        return m, (("obs", m), None)

    builds = []
    real_chunk = functional.CellListPairs.chunk

    def counting_chunk(self, *a, **k):
        builds.append(1)
        return real_chunk(self, *a, **k)

    functional.CellListPairs.chunk = counting_chunk
    old_gb = functional.PAIR_LIST_CACHE_GB
    try:
        functional._PAIR_LISTS.clear()
        with torch.no_grad():
            e_ref = efn.map(states) + obs * 0.01
        n_cold = len(builds)
        assert n_cold >= 1 and functional._PAIR_LISTS.nbytes() > 0
        warm = objective.compute_loss_and_grad(theta, efn, beta, loss_fn, states, e_ref, [])
        assert len(builds) == n_cold, "the second pass must reuse the remembered lists"
        functional.PAIR_LIST_CACHE_GB = 0.0
        cold = objective.compute_loss_and_grad(theta, efn, beta, loss_fn, states, e_ref, [])
        functional.PAIR_LIST_CACHE_GB = old_gb
        assert len(builds) > n_cold
        np.testing.assert_allclose(warm[0][1][2].cpu().numpy(), cold[0][1][2].cpu().numpy(), rtol=1e-12)
        for k in theta:
            np.testing.assert_allclose(float(warm[1][k]), float(cold[1][k]), rtol=1e-9, atol=1e-12)

        # a wider Debye cutoff than the remembered one (lambda_factor +5 % > the 1 % margin) -> rebuilt, and still right
        wide = dict(theta, lambda_factor=theta["lambda_factor"] * 1.05)
        n0 = len(builds)
        with torch.no_grad():
            e_wide = efn.with_params(wide).map(states)
        assert len(builds) > n0
        functional.PAIR_LIST_CACHE_GB = 0.0
        with torch.no_grad():
            e_wide_cold = efn.with_params(wide).map(states)
        functional.PAIR_LIST_CACHE_GB = old_gb
        np.testing.assert_allclose(e_wide.cpu().numpy(), e_wide_cold.cpu().numpy(), rtol=1e-12)
        n0 = len(builds)
        with torch.no_grad():
            efn.map(states)  # narrower cutoffs fit inside the wider remembered lists
        assert len(builds) == n0

        # in-place update of the frames -> version counter moves -> rebuilt
        cc.add_(0.0)
        with torch.no_grad():
            efn.map(states)
        assert len(builds) > n0

        # calculate(): the same trajectory object across calls reaches the remembered lists
        kT = float(dna2.default_configs()[0]["kT"])
        traj = SimulatorTrajectory(center=cc, orientation=Quaternion(qq), temperature=torch.full((F,), kT, dtype=torch.float64, device=DEV))
        objv = objective.DiffTReObjective(name="t", required_observables=("traj",), grad_or_loss_fn=loss_fn, energy_fn=efn,
                                          min_n_eff_factor=0.5, n_equilibration_steps=1)
        out1 = objv.calculate({"traj": traj}, theta)
        n1 = len(builds)
        out2 = objv.calculate({"traj": traj}, theta, opt_steps=1, reference_opt_params=theta)
        assert out1.is_ready and out2.is_ready and len(builds) == n1
        for k in theta:
            np.testing.assert_allclose(float(out1.grads[k]), float(out2.grads[k]), rtol=1e-12, atol=1e-14)
    finally:
        functional.CellListPairs.chunk = real_chunk
        functional.PAIR_LIST_CACHE_GB = old_gb
        functional._PAIR_LISTS.clear()
