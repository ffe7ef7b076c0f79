"""Per-frame observables on the GPU: the standalone kernel and the fused epilogue of the frame-resident
energy kernel against the oracle restatement of mythos/observables/{propeller,rise,pitch,diameter}.py, with the sites
the oracle works on produced by the (torch, host) nucleotide transform -- the route the reference itself takes."""

import numpy as np
import pytest
import torch

from mythos_b200 import observables as obs
from mythos_b200 import space
from mythos_b200.energy import dna1, dna2
from mythos_b200.observables import base as obs_base
from mythos_b200.optimization import objective
from mythos_b200.rigid_body import Quaternion, RigidBody
from mythos_b200.simulators.io import SimulatorTrajectory
from mythos_b200.utils import synthetic
from oracle import observables_oracle as oo

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def duplex_lists(n_dup: int):
    """Base pairs and quartets of `n_dup` 60-bp duplexes laid out as synthetic.assembly does (120 nt per duplex)."""
    q = obs.get_duplex_quartets(60)
    bp = torch.stack([torch.arange(60), torch.arange(119, 59, -1)], 1).to(torch.int32)
    off = (torch.arange(n_dup) * 120).to(torch.int32)
    return (bp[None] + off[:, None, None]).reshape(-1, 2), (q[None] + off[:, None, None, None]).reshape(-1, 2, 2)


def oracle_columns(transform_fn, c, q, bp, qt, sigma, box=None):
    nuc = transform_fn(RigidBody(torch.tensor(c), Quaternion(torch.tensor(q))))
    return oo.frame_columns(nuc.base_sites.numpy(), nuc.back_sites.numpy(), nuc.base_normals.numpy(), bp.numpy(), qt.numpy(), sigma, box)


@pytest.mark.parametrize("model", ["dna1", "dna2"])
@pytest.mark.parametrize("box", [None, 40.0])
def test_standalone_kernel_matches_oracle(model, box):
    s = synthetic.assembly(2, seed=3)
    c, q = synthetic.rejittered_frames(s, 5)
    mod = dna1 if model == "dna1" else dna2
    efn = mod.create_default_energy_fn(s.topology)
    tf = efn.energy_fns[0].transform_fn
    disp = space.periodic(box)[0] if box else space.free()[0]
    bp, qt = duplex_lists(2)
    traj = SimulatorTrajectory(center=torch.tensor(c, device=DEV), orientation=Quaternion(torch.tensor(q, device=DEV)))
    want = oracle_columns(tf, c, q, bp, qt, 0.7, [box] * 3 if box else None)
    got = {
        0: obs.PropellerTwist(rigid_body_transform_fn=tf, h_bonded_base_pairs=bp)(traj),
        1: obs.Rise(rigid_body_transform_fn=tf, quartets=qt, displacement_fn=disp)(traj),
        2: obs.PitchAngle(rigid_body_transform_fn=tf, quartets=qt, displacement_fn=disp)(traj),
        3: obs.Diameter(rigid_body_transform_fn=tf, h_bonded_base_pairs=bp, displacement_fn=disp)(traj, 0.7),
    }
    for k, v in got.items():
        assert v.shape == (5,)
        np.testing.assert_allclose(v.cpu().numpy(), want[:, k], rtol=1e-10, err_msg=f"column {k}")
    # sanity of the synthetic B-form duplex against the reference's targets (loose: ideal geometry + jitter)
    # (the synthetic duplex has exactly antiparallel base normals before jitter, so its propeller twist is only the jitter)
    assert 0.0 < float(got[0].mean()) < 30.0 and 3.0 < float(got[1].mean()) < 3.8 and 20.0 < float(got[3].mean()) < 26.0
    assert 0.55 < float(got[2].mean()) < 0.70  # the construction's 35.9 degree twist per base pair = 0.627 rad
    # float32: the north star's 1e-4
    traj32 = SimulatorTrajectory(center=traj.center.float(), orientation=Quaternion(traj.orientation.vec.float()))
    r32 = obs.Rise(rigid_body_transform_fn=tf, quartets=qt, displacement_fn=disp)(traj32)
    np.testing.assert_allclose(r32.cpu().numpy(), want[:, 1], rtol=1e-4)


def test_empty_lists_give_nan_like_jnp_mean():
    s = synthetic.assembly(1, seed=3)
    efn = dna2.create_default_energy_fn(s.topology)
    tf = efn.energy_fns[0].transform_fn
    body = RigidBody(torch.tensor(s.center, device=DEV)[None], Quaternion(torch.tensor(s.quat, device=DEV)[None]))
    cols = obs_base.columns(tf, space.free()[0], body, base_pairs=torch.tensor([[0, 119]]), quartets=None)
    assert torch.isfinite(cols[:, 0]).all() and torch.isnan(cols[:, 1]).all() and torch.isnan(cols[:, 2]).all()


def test_fused_epilogue_of_the_energy_pass_equals_standalone_and_feeds_the_loss():
    """``map(states, observables=ObservableSet)`` evaluates the observables in the frame-resident kernel's epilogue; the member
    observables then answer from that result without a launch; ``compute_loss`` does the same for
    ``loss_fn.fused_observables``; energies are untouched."""
    s = synthetic.assembly(17, seed=1)
    c, q = synthetic.rejittered_frames(s, 6)
    efn = dna2.create_default_energy_fn(s.topology)
    tf = efn.energy_fns[0].transform_fn
    disp = efn.energy_fns[0].displacement_fn
    bp, qt = duplex_lists(17)
    sigma = float(efn.params_dict()["sigma_backbone"])
    members = (obs.PropellerTwist(rigid_body_transform_fn=tf, h_bonded_base_pairs=bp),
               obs.Rise(rigid_body_transform_fn=tf, quartets=qt, displacement_fn=disp),
               obs.PitchAngle(rigid_body_transform_fn=tf, quartets=qt, displacement_fn=disp),
               obs.Diameter(rigid_body_transform_fn=tf, h_bonded_base_pairs=bp, displacement_fn=disp, sigma_backbone=sigma))
    oset = obs.ObservableSet(members)
    want = oracle_columns(tf, c, q, bp, qt, sigma)

    launches = []
    real = obs_base.launch
    obs_base.launch = lambda *a, **k: (launches.append(1), real(*a, **k))[1]
    try:
        cc, qq = torch.tensor(c, device=DEV), torch.tensor(q, device=DEV)
        states = SimulatorTrajectory(center=cc, orientation=Quaternion(qq), temperature=torch.full((6,), 0.1, dtype=torch.float64, device=DEV))
        with torch.no_grad():
            e_plain = efn.map(RigidBody(cc.clone(), Quaternion(qq.clone())))
            e_fused = efn.map(states, observables=oset)
        np.testing.assert_allclose(e_fused.cpu().numpy(), e_plain.cpu().numpy(), rtol=1e-13)
        fused = [members[0](states), members[1](states), members[2](states), members[3](states)]
        assert not launches, "the members must answer from the fused result"
        for k, v in enumerate(fused):
            np.testing.assert_allclose(v.cpu().numpy(), want[:, k], rtol=1e-10, err_msg=f"column {k}")
        alone = members[1](SimulatorTrajectory(center=cc.clone(), orientation=Quaternion(qq.clone())))
        assert launches == [1]
        np.testing.assert_allclose(alone.cpu().numpy(), fused[1].cpu().numpy(), rtol=1e-13)

        # compute_loss: the loss function declares its observables; its own observable(ref_states) call costs nothing
        obs_base.COLUMNS.clear()
        launches.clear()
        target = 21.7

        def loss_fn(ref_states, weights, energy_fn, opt_params, observables):
            expected = (weights * members[0](ref_states)).sum()
            return (expected - target) ** 2, (("propeller", expected), None)

        loss_fn.fused_observables = oset
        theta = {"eps_hb": torch.tensor(float(efn.params_dict()["eps_hb"]), dtype=torch.float64)}
        beta = torch.full((6,), 10.0, dtype=torch.float64, device=DEV)
        (loss, (neff, measured, _)), grads = objective.compute_loss_and_grad(theta, efn, beta, loss_fn, states, e_plain + 0.01, [])
        assert not launches
        w = torch.softmax(-beta * (-0.01), 0)
        np.testing.assert_allclose(float(measured[1]), float((w * torch.tensor(want[:, 0], device=DEV)).sum()), rtol=1e-9)
        assert np.isfinite(float(grads["eps_hb"]))
    finally:
        obs_base.launch = real
        obs_base.COLUMNS.clear()


def test_observable_set_refuses_mismatched_members():
    s = synthetic.assembly(1, seed=3)
    efn = dna2.create_default_energy_fn(s.topology)
    tf = efn.energy_fns[0].transform_fn
    a = obs.PropellerTwist(rigid_body_transform_fn=tf, h_bonded_base_pairs=torch.tensor([[0, 119]]))
    b = obs.Diameter(rigid_body_transform_fn=tf, h_bonded_base_pairs=torch.tensor([[1, 118]]), displacement_fn=space.free()[0], sigma_backbone=0.7)
    with pytest.raises(ValueError):
        obs.ObservableSet([a, b]).request(torch.device(DEV))
