"""Differentiable-through-time MD: gradient of a trajectory loss with respect to theta and the initial state
(SURVEY 8f rank 4).

The reference differentiates through ``lax.scan`` of ``step_fn`` with ``jax.checkpoint`` every few hundred steps
(``mythos/simulators/jax_md/utils.py:174-193``, ``jaxmd.py:54-58,94``): reverse mode through the integrator AND through the
energy function's own gradient, i.e. second derivatives of every term by autodiff.  Here the backward pass is explicit:

* the forward run stores, per step, the pre-step state, the forces it used and the noise (N x 27 reals per step);
* ``mythos_b200_langevin_adjoint_*`` is the hand-written vector-Jacobian product of the integrator sub-steps
  (kick, free-rotor drifts, Ornstein-Uhlenbeck), with the forces as an input;
* the force's own dependence on positions and parameters -- the Hessian-vector product ``H v`` and the mixed derivative
  ``d(v . dE/dx)/dparams`` for the cotangent ``v`` of the forces -- comes from TWO evaluations of the analytic
  energy + force + dE/dparams kernel at positions displaced by ``+-eps v`` (central difference of analytic first
  derivatives: in float64 the truncation and round-off errors are both ~1e-10 relative at the default displacement of 1e-5
  length units; tests hold it to 1e-6 against autograd through the oracle).
"""

from __future__ import annotations

import ctypes as C

import torch

from mythos_b200 import _lib
from mythos_b200.energy import functional
from mythos_b200.rigid_body import Quaternion, RigidBody
from mythos_b200.simulators import md
from mythos_b200.simulators.io import SimulatorTrajectory


def _scalar(x) -> float:
    return float(torch.as_tensor(x).reshape(-1)[0])


class _Stepper:
    """Raw C-ABI calls of the forward step and its adjoint for one set of integrator constants."""

    def __init__(self, dt, kT, gamma, mass, dtype):
        self.dt, self.kT = float(dt), float(kT)
        if isinstance(gamma, RigidBody):
            self.gamma_c, self.gamma_q = _scalar(gamma.center), _scalar(gamma.orientation)
        else:
            self.gamma_c = self.gamma_q = float(gamma)
        mass = mass if mass is not None else RigidBody(torch.tensor(1.0), torch.tensor([1.0, 1.0, 1.0]))
        self.mass = _scalar(mass.center)
        self.inertia = [float(v) for v in torch.as_tensor(mass.orientation, dtype=torch.float64).reshape(-1)[:3]]
        self.sfx = _lib.suffix(dtype)

    def forward(self, c, q, pc, pq, g_c, g_q, noise, phase: int) -> None:
        a = _lib.LangevinArgs()
        a.n = c.shape[0]
        a.center, a.quat, a.p_center, a.p_quat = c.data_ptr(), q.data_ptr(), pc.data_ptr(), pq.data_ptr()
        a.d_center, a.d_quat = g_c.data_ptr(), g_q.data_ptr()
        a.dt, a.kT, a.gamma_center, a.gamma_quat, a.mass = self.dt, self.kT, self.gamma_c, self.gamma_q, self.mass
        for d in range(3):
            a.inertia[d] = self.inertia[d]
            a.box[d] = 0.0
        a.noise = noise.data_ptr()
        a.phase = phase
        with torch.cuda.device(c.device):
            _lib.check(getattr(_lib.lib(), f"mythos_b200_langevin_{self.sfx}")(_lib.current_stream(c.device), C.byref(a)), "mythos_b200_langevin")

    def adjoint(self, c, q, pc, pq, g_c, g_q, noise, phase: int, lam, lam_force) -> None:
        a = _lib.LangevinAdjointArgs()
        a.n, a.phase = c.shape[0], phase
        a.center, a.quat, a.p_center, a.p_quat = c.data_ptr(), q.data_ptr(), pc.data_ptr(), pq.data_ptr()
        a.d_center, a.d_quat, a.noise = g_c.data_ptr(), g_q.data_ptr(), noise.data_ptr()
        a.lam_center, a.lam_quat, a.lam_p_center, a.lam_p_quat = (t.data_ptr() for t in lam)
        a.lam_force_center, a.lam_force_quat = (t.data_ptr() for t in lam_force)
        a.dt, a.kT, a.gamma_center, a.gamma_quat, a.mass = self.dt, self.kT, self.gamma_c, self.gamma_q, self.mass
        for d in range(3):
            a.inertia[d] = self.inertia[d]
        with torch.cuda.device(c.device):
            _lib.check(getattr(_lib.lib(), f"mythos_b200_langevin_adjoint_{self.sfx}")(_lib.current_stream(c.device), C.byref(a)),
                       "mythos_b200_langevin_adjoint")


def simulate_and_grad(energy_fn, opt_params: dict, init_state: RigidBody, n_steps: int, loss_fn, *, dt: float, kT: float, gamma,
                      mass: RigidBody | None = None, key: int = 0, init_momentum: RigidBody | None = None, noise: torch.Tensor | None = None,
                      fd_displacement: float = 1e-5):
    """Run ``n_steps`` of fused Langevin MD (free space, static unbonded list) and differentiate
    ``loss_fn(trajectory: SimulatorTrajectory) -> scalar`` through the whole trajectory.

    Returns ``(loss, grads, trajectory, init_grad)``: ``grads[name] = dloss/dopt_params[name]``, ``init_grad`` the
    RigidBody of gradients with respect to the initial centres / quaternions.  ``noise (n_steps, N, 6)`` fixes the thermal
    noise (default: drawn from ``key``); ``init_momentum`` the initial momenta (default: zero)."""
    _lib.require_cuda(init_state.center, "init_state.center")
    dev, dtype = init_state.center.device, init_state.center.dtype
    n = init_state.center.shape[0]
    leaves = {k: torch.as_tensor(v, dtype=torch.float64).detach().clone().requires_grad_(True) for k, v in opt_params.items()}
    efn = energy_fn.with_params(leaves) if leaves else energy_fn
    plan, cot = md._plan_of(efn)
    topo = plan.topology(n, dev)
    bank_host = plan.params_vector()
    params = bank_host.detach().to(device=dev, dtype=dtype)
    source = plan.pairs(dev, topo)
    if not isinstance(source, functional.StaticPairs):
        raise _lib.MythosB200Error("simulate_and_grad needs an explicit unbonded pair list (the reference's default topology list)")
    cot_dev = cot.to(device=dev, dtype=dtype).reshape(1, -1).contiguous()

    def gradients(c, q, want_params: bool):
        _, dc, dq, dp = functional.energy_and_gradients(plan.model, topo, c.unsqueeze(0), q.unsqueeze(0), params, source, cot=cot_dev,
                                                        want_pos_grad=True, want_param_grad=want_params)
        return dc[0], dq[0], dp

    stepper = _Stepper(dt, kT, gamma, mass, dtype)
    if noise is None:
        gen = torch.Generator(device=dev)
        gen.manual_seed(int(key))
        noise = torch.randn((n_steps, n, 6), generator=gen, device=dev, dtype=dtype)
    noise = noise.to(device=dev, dtype=dtype).contiguous()
    c, q = init_state.center.detach().clone(), init_state.orientation.vec.detach().clone()
    pc = torch.zeros((n, 3), device=dev, dtype=dtype) if init_momentum is None else init_momentum.center.detach().clone()
    pq = torch.zeros((n, 4), device=dev, dtype=dtype) if init_momentum is None else init_momentum.orientation.vec.detach().clone()

    # ---- forward: pre-step state, forces and noise of every step are kept (N x 27 reals per step)
    saved = []
    traj_c = torch.empty((n_steps, n, 3), device=dev, dtype=dtype)
    traj_q = torch.empty((n_steps, n, 4), device=dev, dtype=dtype)
    for k in range(n_steps):
        g_c, g_q, _ = gradients(c, q, False)
        saved.append((c.clone(), q.clone(), pc.clone(), pq.clone(), g_c, g_q))
        stepper.forward(c, q, pc, pq, g_c, g_q, noise[k], 0 if k == 0 else 2)
        traj_c[k], traj_q[k] = c, q

    # ---- loss on the trajectory (torch autograd gives its gradient with respect to every stored position)
    tc, tq = traj_c.clone().requires_grad_(True), traj_q.clone().requires_grad_(True)
    states = SimulatorTrajectory(center=tc, orientation=Quaternion(tq), temperature=torch.full((n_steps,), float(kT), device=dev, dtype=dtype))
    loss = loss_fn(states)
    g_tc, g_tq = torch.autograd.grad(loss, [tc, tq], allow_unused=True)
    g_tc = torch.zeros_like(tc) if g_tc is None else g_tc
    g_tq = torch.zeros_like(tq) if g_tq is None else g_tq

    # ---- backward sweep
    lam = [torch.zeros((n, 3), device=dev, dtype=dtype), torch.zeros((n, 4), device=dev, dtype=dtype),
           torch.zeros((n, 3), device=dev, dtype=dtype), torch.zeros((n, 4), device=dev, dtype=dtype)]
    lam_force = [torch.empty((n, 3), device=dev, dtype=dtype), torch.empty((n, 4), device=dev, dtype=dtype)]
    g_bank = torch.zeros_like(params)
    for k in range(n_steps - 1, -1, -1):
        lam[0] += g_tc[k]
        lam[1] += g_tq[k]
        c0, q0, pc0, pq0, g_c, g_q = saved[k]
        stepper.adjoint(c0, q0, pc0, pq0, g_c, g_q, noise[k], 0 if k == 0 else 2, lam, lam_force)
        # forces depend on (x_k, theta): H v and d(v . dE/dx)/dparams by a central difference of analytic gradients along v
        scale = torch.maximum(lam_force[0].abs().max(), lam_force[1].abs().max())
        if float(scale) > 0.0:  # (one host read per step: the displacement is normalised to a fixed length)
            eps = fd_displacement / float(scale)
            dc_p, dq_p, dp_p = gradients(c0 + eps * lam_force[0], q0 + eps * lam_force[1], True)
            dc_m, dq_m, dp_m = gradients(c0 - eps * lam_force[0], q0 - eps * lam_force[1], True)
            lam[0] += (dc_p - dc_m) / (2 * eps)
            lam[1] += (dq_p - dq_m) / (2 * eps)
            g_bank += (dp_p - dp_m) / (2 * eps)
    grads = {}
    if leaves:
        gl = torch.autograd.grad(bank_host, list(leaves.values()), grad_outputs=g_bank.to(device=bank_host.device, dtype=bank_host.dtype),
                                 allow_unused=True)
        grads = {k: (torch.zeros_like(v) if g is None else g) for (k, v), g in zip(leaves.items(), gl)}
    traj = SimulatorTrajectory(center=traj_c, orientation=Quaternion(traj_q), temperature=torch.full((n_steps,), float(kT), device=dev, dtype=dtype))
    return loss.detach(), grads, traj, RigidBody(lam[0], Quaternion(lam[1]))
