"""CPU restatement of the rigid-body Langevin (BAOAB) step.  TEST INFRASTRUCTURE ONLY.

The algorithm lives in jax_md==0.2.28 (simulate.nvt_langevin + rigid_body dispatch), a third-party dependency that
is NOT vendored in /root/reference and cannot be installed here; the reference's own tests use a fake integrator
(mythos/simulators/jax_md/tests/test_jaxmd.py:100-124).  PARITY UNPINNED: this file restates the published scheme
as recalled in SURVEY appendix D (BAOAB; quaternion free-rotor splitting of Miller et al., J. Chem. Phys. 116, 8649
(2002), rotation order 3,2,1,2,3; Ornstein-Uhlenbeck on the body-frame angular momentum) in plain numpy, so the CUDA
kernel can be checked sub-step by sub-step with injected noise, plus statistical checks (equipartition).
"""

from __future__ import annotations

import numpy as np


def perm(k: int, q: np.ndarray) -> np.ndarray:
    """P_k q = column k of S(q) (Miller et al. eq. 2.12-2.14)."""
    q0, q1, q2, q3 = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    if k == 1:
        return np.stack([-q1, q0, q3, -q2], -1)
    if k == 2:
        return np.stack([-q2, -q3, q0, q1], -1)
    return np.stack([-q3, q2, -q1, q0], -1)


def free_rotor(k, step, inertia_k, q, p):
    pq, pp = perm(k, q), perm(k, p)
    zeta = step * (p * pq).sum(-1, keepdims=True) / (4.0 * inertia_k)
    c, s = np.cos(zeta), np.sin(zeta)
    return c * q + s * pq, c * p + s * pp


def drift(c, q, pc, pq, h, mass, inertia, box=None):
    c = c + h * pc / mass
    if box is not None:
        c = np.mod(c, np.asarray(box))
    for k, st in ((3, 0.5 * h), (2, 0.5 * h), (1, h), (2, 0.5 * h), (3, 0.5 * h)):
        q, pq = free_rotor(k, st, inertia[k - 1], q, pq)
    return c, q, pq


def angular_momentum(q, pq):
    return np.stack([0.5 * (pq * perm(k, q)).sum(-1) for k in (1, 2, 3)], -1)


def conjugate_momentum(q, L):
    return 2.0 * sum(L[..., k - 1 : k] * perm(k, q) for k in (1, 2, 3))


def step(c, q, pc, pq, d_center, d_quat, noise, dt, kT, gamma_c, gamma_q, mass, inertia, box=None, kick=None):
    """B(kick) A(dt/2) O(dt) A(dt/2); ``kick`` defaults to dt/2 (phase 0), dt for the fused phase 2."""
    h = 0.5 * dt
    kick = h if kick is None else kick
    pc = pc - kick * d_center
    pq = pq - kick * d_quat
    c, q, pq = drift(c, q, pc, pq, h, mass, inertia, box)
    c1 = np.exp(-gamma_c * dt)
    pc = c1 * pc + np.sqrt(kT * (1 - c1 * c1) * mass) * noise[:, :3]
    r1 = np.exp(-gamma_q * dt)
    L = r1 * angular_momentum(q, pq) + np.sqrt(kT * (1 - r1 * r1) * np.asarray(inertia)) * noise[:, 3:]
    pq = conjugate_momentum(q, L)
    c, q, pq = drift(c, q, pc, pq, h, mass, inertia, box)
    return c, q, pc, pq


def kinetic_energies(q, pc, pq, mass, inertia):
    L = angular_momentum(q, pq)
    return (pc * pc).sum(-1) / (2 * mass), (L * L / (2 * np.asarray(inertia))).sum(-1)


# ---------------------------------------------------------------------------------------------------------------------
# The same step on torch tensors, so that torch.autograd can differentiate THROUGH it: the oracle for the adjoint of the
# step (what jax.grad does through step_fn inside checkpoint_scan, mythos/simulators/jax_md/utils.py:174-193).
def _perm_t(k, q):
    import torch

    q0, q1, q2, q3 = q.unbind(-1)
    if k == 1:
        return torch.stack([-q1, q0, q3, -q2], -1)
    if k == 2:
        return torch.stack([-q2, -q3, q0, q1], -1)
    return torch.stack([-q3, q2, -q1, q0], -1)


def _drift_t(c, q, pc, pq, h, mass, inertia):
    import torch

    c = c + h * pc / mass
    for k, st in ((3, 0.5 * h), (2, 0.5 * h), (1, h), (2, 0.5 * h), (3, 0.5 * h)):
        Pq, Pp = _perm_t(k, q), _perm_t(k, pq)
        zeta = st * (pq * Pq).sum(-1, keepdim=True) / (4.0 * inertia[k - 1])
        cz, sz = torch.cos(zeta), torch.sin(zeta)
        q, pq = cz * q + sz * Pq, cz * pq + sz * Pp
    return c, q, pq


def step_torch(c, q, pc, pq, d_center, d_quat, noise, dt, kT, gamma_c, gamma_q, mass, inertia, kick=None):
    """``step`` with torch tensors (free space), differentiable in every tensor argument."""
    import math

    import torch

    h = 0.5 * dt
    kick = h if kick is None else kick
    pc = pc - kick * d_center
    pq = pq - kick * d_quat
    c, q, pq = _drift_t(c, q, pc, pq, h, mass, inertia)
    c1 = math.exp(-gamma_c * dt)
    pc = c1 * pc + math.sqrt(kT * (1 - c1 * c1) * mass) * noise[:, :3]
    r1 = math.exp(-gamma_q * dt)
    L = torch.stack([0.5 * (pq * _perm_t(k, q)).sum(-1) for k in (1, 2, 3)], -1)
    L = r1 * L + torch.sqrt(kT * (1 - r1 * r1) * torch.as_tensor(inertia, dtype=c.dtype)) * noise[:, 3:]
    pq = 2.0 * sum(L[..., k - 1 : k] * _perm_t(k, q) for k in (1, 2, 3))
    c, q, pq = _drift_t(c, q, pc, pq, h, mass, inertia)
    return c, q, pc, pq
