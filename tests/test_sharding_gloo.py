"""Host-side logic of the frame-sharded DiffTRe pass, world_size 2 over gloo on CPU.

The energy kernels need a GPU, so the energy function here is a CPU test double with the EnergyFunction surface
(with_params / map); what is under test is the sharding: contiguous blocks, all-gather of the per-frame energies,
slice-only backward, all-reduce of the frame part of the gradient plus the replicated direct part.
"""

import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mythos_b200.optimization import objective
from mythos_b200.rigid_body import Quaternion, RigidBody


class FakeEnergy:
    """E_k = a * sum(center_k^2) + b * sum(center_k) -- linear in theta, per-frame, CPU."""

    def __init__(self, theta=None):
        self.theta = theta or {"a": torch.tensor(0.7, dtype=torch.float64), "b": torch.tensor(-0.2, dtype=torch.float64)}

    def with_params(self, theta):
        return FakeEnergy({**self.theta, **theta})

    def map(self, states):
        c = states.center
        return self.theta["a"] * (c * c).sum((1, 2)) + self.theta["b"] * c.sum((1, 2))


def _cpu_weights(beta, new_e, ref_e):
    d = new_e - ref_e
    w = torch.softmax(-beta * d, 0)
    return w, torch.exp(-(w * torch.log(w)).sum()) / len(w)


def _loss_fn(ref_states, weights, energy_fn, opt_params, observables):
    obs = observables[0]
    m = (weights * obs).sum()
    # direct theta dependence too: through opt_params and through the re-parameterised energy function
    return (m - 0.3) ** 2 + 0.01 * opt_params["a"] ** 2 + 0.005 * energy_fn.theta["b"] ** 2, (("obs", m), None)


def _worker(rank, world, port, out, F=11):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        objective.compute_weights_and_neff = _cpu_weights  # the CUDA reweighting kernel is not under test here
        g = torch.Generator().manual_seed(0)
        c = torch.randn((F, 5, 3), generator=g, dtype=torch.float64)
        q = torch.randn((F, 5, 4), generator=g, dtype=torch.float64)
        obs = torch.randn(F, generator=g, dtype=torch.float64)
        states = RigidBody(c, Quaternion(q))
        efn = FakeEnergy()
        theta = {"a": torch.tensor(0.75, dtype=torch.float64), "b": torch.tensor(-0.1, dtype=torch.float64)}
        ref = efn.map(states).detach()
        (loss, aux), grads = objective.compute_loss_and_grad(theta, efn, torch.tensor(2.0, dtype=torch.float64), _loss_fn, states, ref, [obs])
        lo, hi = objective.shard_bounds(F, rank, world)
        out[rank] = (float(loss), {k: float(v) for k, v in grads.items()}, (lo, hi), aux[2].numpy().copy())
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_all_frames():
    for F in (1, 7, 8, 8192, 8193):
        for w in (1, 2, 3, 8):
            blocks = [objective.shard_bounds(F, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == F
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(120)
@pytest.mark.parametrize("F", [11, 12])  # blocks of 6 + 5 frames (padded gather) and 6 + 6 (single-collective gather)
def test_two_rank_gradient_equals_single_process(F):
    port = 29500 + (os.getpid() % 1000) + F
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(2, port, out, F), nprocs=2, join=True)
        res = dict(out)
    # single-process answer
    g = torch.Generator().manual_seed(0)
    c = torch.randn((F, 5, 3), generator=g, dtype=torch.float64)
    q = torch.randn((F, 5, 4), generator=g, dtype=torch.float64)
    obs = torch.randn(F, generator=g, dtype=torch.float64)
    states = RigidBody(c, Quaternion(q))
    efn = FakeEnergy()
    theta = {"a": torch.tensor(0.75, dtype=torch.float64, requires_grad=True), "b": torch.tensor(-0.1, dtype=torch.float64, requires_grad=True)}
    ref = efn.map(states).detach()
    efn_theta = efn.with_params(theta)
    e = efn_theta.map(states)
    w, _ = _cpu_weights(torch.tensor(2.0, dtype=torch.float64), e, ref)
    loss, _ = _loss_fn(states, w, efn_theta, theta, [obs])
    ga, gb = torch.autograd.grad(loss, [theta["a"], theta["b"]])
    for rank in (0, 1):
        l, grads, bounds, e_full = res[rank]
        assert np.isclose(l, float(loss), rtol=1e-12)
        assert np.isclose(grads["a"], float(ga), rtol=1e-10) and np.isclose(grads["b"], float(gb), rtol=1e-10)
        np.testing.assert_allclose(e_full, e.detach().numpy(), rtol=1e-13)
    assert res[0][2] == (0, 6) and res[1][2] == (6, F)
