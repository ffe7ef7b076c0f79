"""In-process MD on the device: fused rigid-body Langevin integrator + run loop.

Interfaces kept from the reference:

* ``nvt_langevin(energy_fn, shift_fn, dt, kT, gamma)`` has the calling convention of the ``simulator_init`` plug-in
  that ``mythos/simulators/jax_md/jaxmd.py:27,73`` receives (the user passes ``jax_md.simulate.nvt_langevin``):
  it returns ``(init_fn, step_fn)`` with ``init_fn(key, R, mass=..., unbonded_neighbors=...)`` -> state holding
  ``.position`` and ``.mass``, and ``step_fn(state, unbonded_neighbors=...)`` -> state.
* ``MDSimulator.run(opt_params, init_state, n_steps, key)`` mirrors ``build_run_fn`` (``jaxmd.py:60-101``): parameters
  are applied once, every step's position is emitted, the result is a ``SimulatorTrajectory`` with per-state kT.

Underneath, one step is two launches: the fused B-A-O-A kernel (``csrc/langevin.cu``: half kicks, drift, free-rotor
quaternion update, Ornstein-Uhlenbeck, trajectory row store, counter-based RNG) and the fused energy+force pair
kernels; ``run`` captures the step in a CUDA graph and replays it, the step index living in a device counter.
"""

from __future__ import annotations

import ctypes as C
import dataclasses as dc
from typing import Any

import torch

from mythos_b200 import _lib, space
from mythos_b200.energy import functional
from mythos_b200.energy import model as kmodel
from mythos_b200.rigid_body import Quaternion, RigidBody
from mythos_b200.simulators.io import SimulatorTrajectory


@dc.dataclass
class NVTLangevinState:
    position: RigidBody
    momentum: RigidBody  # center: linear momentum (N,3); orientation.vec: quaternion-conjugate momentum (N,4)
    force: RigidBody  # holds dE/dcenter, dE/dquat (the force is its negative)
    mass: RigidBody  # center: scalar mass; orientation: (3,) principal moments
    step: torch.Tensor  # (2,) uint64-as-int64: [0] device step counter, [1] kernel scratch
    seed: int = 0


def _perm(k: int, q: torch.Tensor) -> torch.Tensor:
    q0, q1, q2, q3 = q.unbind(-1)
    if k == 1:
        return torch.stack([-q1, q0, q3, -q2], -1)
    if k == 2:
        return torch.stack([-q2, -q3, q0, q1], -1)
    return torch.stack([-q3, q2, -q1, q0], -1)


def _plan_of(energy_fn):
    fns = energy_fn.energy_fns if hasattr(energy_fn, "energy_fns") else [energy_fn]
    groups = kmodel.fusable_groups(fns)
    if len(groups) != 1:
        raise _lib.MythosB200Error("the fused integrator needs an energy function whose terms share one launch group")
    weights = torch.ones(len(fns), dtype=torch.float64) if getattr(energy_fn, "weights", None) is None else torch.as_tensor(energy_fn.weights, dtype=torch.float64)
    cot = torch.zeros(_lib.N_TERMS, dtype=torch.float64)
    for k, fn in enumerate(fns):
        cot[fn.TERM] = weights[k]
    return kmodel.plan_for(fns), cot


class _Forces:
    """Energy + (dE/dcenter, dE/dquat) of the current positions through the C-ABI, buffers reused across steps."""

    def __init__(self, energy_fn, n: int, device, dtype, unbonded_neighbors=None):
        if unbonded_neighbors is not None:
            energy_fn = energy_fn.with_props(unbonded_neighbors=unbonded_neighbors)
        self.plan, cot = _plan_of(energy_fn)
        self.topo = self.plan.topology(n, device)
        self.params = self.plan.device_params(device, dtype).detach()
        self.source = self.plan.pairs(device, self.topo)
        self.cot = cot.to(device=device, dtype=dtype).reshape(1, -1).contiguous()

    def __call__(self, center: torch.Tensor, quat: torch.Tensor):
        terms, dcen, dq, _ = functional.energy_and_gradients(
            self.plan.model, self.topo, center.unsqueeze(0), quat.unsqueeze(0), self.params, self.source, cot=self.cot,
            want_pos_grad=True)
        return terms[0], dcen[0], dq[0]

    def accumulate_into(self, center: torch.Tensor, quat: torch.Tensor, d_center: torch.Tensor, d_quat: torch.Tensor) -> None:
        """Forces only, ADDED into caller-owned (already zeroed) buffers: two pair-kernel launches, no memsets, no energies."""
        if not isinstance(self.source, functional.StaticPairs):
            raise _lib.MythosB200Error("accumulate_into needs an explicit pair list")
        pairs, stride, _ = self.source.chunk(slice(0, 1), center)
        functional._launch(self.plan.model, self.topo, center.unsqueeze(0), quat.unsqueeze(0), self.params, pairs, stride,
                           self.plan.term_mask, self.cot, False, True, False, False, None, _lib.FLAG_ACCUMULATE, 0.0,
                           (d_center.unsqueeze(0), d_quat.unsqueeze(0)))


def nvt_langevin(energy_fn, shift_fn, dt: float, kT: float, gamma: RigidBody | Any = 0.1, seed: int = 0):
    """Fused rigid-body BAOAB Langevin integrator with the ``simulator_init`` calling convention."""
    box = shift_fn.box if getattr(shift_fn, "box", None) is not None else (0.0, 0.0, 0.0)
    if isinstance(gamma, RigidBody):
        gamma_c, gamma_q = float(gamma.center), float(torch.as_tensor(gamma.orientation).reshape(-1)[0])
    else:
        gamma_c = gamma_q = float(gamma)
    cache: dict[str, Any] = {}

    def forces_for(R: RigidBody, unbonded_neighbors):
        key = id(unbonded_neighbors)
        if cache.get("key") != key:
            cache["key"] = key
            cache["f"] = _Forces(energy_fn, R.center.shape[0], R.center.device, R.center.dtype, unbonded_neighbors)
        return cache["f"]

    def launch(state: NVTLangevinState, phase: int, noise=None, traj=None, advance=False, zero_forces=False):
        c, q = state.position.center, state.position.orientation.vec
        a = _lib.LangevinArgs()
        a.n = c.shape[0]
        a.center, a.quat = c.data_ptr(), q.data_ptr()
        a.p_center, a.p_quat = state.momentum.center.data_ptr(), state.momentum.orientation.vec.data_ptr()
        a.d_center, a.d_quat = state.force.center.data_ptr(), state.force.orientation.vec.data_ptr()
        a.dt, a.kT, a.gamma_center, a.gamma_quat = float(dt), float(kT), gamma_c, gamma_q
        a.mass = float(torch.as_tensor(state.mass.center).reshape(-1)[0])
        inertia = torch.as_tensor(state.mass.orientation, dtype=torch.float64).reshape(-1)
        for d in range(3):
            a.inertia[d] = float(inertia[d])
            a.box[d] = float(box[d])
        a.seed, a.step = state.seed, 0
        a.noise = _lib.ptr(noise)
        a.phase = phase
        a.advance_step = 1 if advance else 0
        a.zero_forces = 1 if zero_forces else 0
        a.step_ptr = state.step.data_ptr()
        if traj is not None:
            a.traj_center, a.traj_quat, a.traj_rows = traj[0].data_ptr(), traj[1].data_ptr(), traj[0].shape[0]
        fn = getattr(_lib.lib(), f"mythos_b200_langevin_{_lib.suffix(c.dtype)}")
        with torch.cuda.device(c.device):
            _lib.check(fn(_lib.current_stream(c.device), C.byref(a)), "mythos_b200_langevin")

    def init_fn(key, R: RigidBody, mass: RigidBody | None = None, unbonded_neighbors=None, **_kw) -> NVTLangevinState:
        """Maxwell-Boltzmann momenta at kT (centre-of-mass momentum removed), forces at R."""
        _lib.require_cuda(R.center, "R.center")
        dev, dtype = R.center.device, R.center.dtype
        if mass is None:
            mass = RigidBody(torch.tensor(1.0), torch.tensor([1.0, 1.0, 1.0]))
        gen = torch.Generator(device=dev)
        gen.manual_seed(int(key))
        m = float(torch.as_tensor(mass.center).reshape(-1)[0])
        inertia = torch.as_tensor(mass.orientation, dtype=dtype, device=dev).reshape(-1)
        n = R.center.shape[0]
        pc = (m * kT) ** 0.5 * torch.randn((n, 3), generator=gen, device=dev, dtype=dtype)
        pc = pc - pc.mean(0, keepdim=True)
        L = torch.sqrt(inertia * kT) * torch.randn((n, 3), generator=gen, device=dev, dtype=dtype)
        q = R.orientation.vec
        pq = 2.0 * sum(L[:, k - 1 : k] * _perm(k, q) for k in (1, 2, 3))
        f = forces_for(R, unbonded_neighbors)
        _, dcen, dq = f(R.center, q)
        return NVTLangevinState(
            position=RigidBody(R.center.clone(), Quaternion(q.clone())),
            momentum=RigidBody(pc.contiguous(), Quaternion(pq.contiguous())),
            force=RigidBody(dcen, Quaternion(dq)),
            mass=mass,
            step=torch.zeros(2, dtype=torch.int64, device=dev),  # [0] step counter, [1] scratch of the kernel
            seed=int(key) + seed,
        )

    def step_fn(state: NVTLangevinState, unbonded_neighbors=None, noise=None, traj=None, **_kw) -> NVTLangevinState:
        """One full BAOAB step in place: B A O A (kernel), force evaluation, closing B (kernel)."""
        launch(state, 0, noise=noise, traj=traj, advance=True)
        f = forces_for(state.position, unbonded_neighbors)
        _, dcen, dq = f(state.position.center, state.position.orientation.vec)
        state.force.center.copy_(dcen)
        state.force.orientation.vec.copy_(dq)
        launch(state, 1)
        return state

    step_fn.launch = launch
    step_fn.forces_for = forces_for
    return init_fn, step_fn


@dc.dataclass
class StaticSimulatorParams:
    """``mythos/simulators/jax_md/utils.py:129-159``"""

    seq: Any
    mass: RigidBody
    gamma: RigidBody
    bonded_neighbors: Any
    checkpoint_every: int
    dt: float
    kT: float  # noqa: N815

    @property
    def sim_init_fn(self) -> dict:
        return {"dt": self.dt, "kT": self.kT, "gamma": self.gamma}

    @property
    def init_fn(self) -> dict:
        return {"mass": self.mass}

    @property
    def step_fn(self) -> dict:
        return {}


@dc.dataclass
class NoNeighborList:
    """Static unbonded list (``simulators/jax_md/utils.py:49-67``)."""

    unbonded_nbrs: Any

    @property
    def idx(self):
        return self.unbonded_nbrs

    def allocate(self, locs):
        return self

    def update(self, locs):
        return self


@dc.dataclass
class MDSimulator:
    """In-process differentiable-state MD runner with the interface of ``JaxMDSimulator`` (``jaxmd.py:21-103``)."""

    energy_fn: Any
    simulator_params: StaticSimulatorParams
    space: tuple  # (displacement_fn, shift_fn)
    simulator_init: Any = nvt_langevin
    neighbors: Any = None
    use_cuda_graph: bool = True

    def run(self, opt_params: dict, init_state: RigidBody, n_steps: int, key: int = 0) -> SimulatorTrajectory:
        _, shift_fn = self.space
        efn = self.energy_fn.with_params(opt_params) if opt_params else self.energy_fn
        neighbors = self.neighbors
        if neighbors is None:
            fns = efn.energy_fns if hasattr(efn, "energy_fns") else [efn]
            neighbors = NoNeighborList(unbonded_nbrs=next(fn.unbonded_neighbors for fn in fns if fn.TERM >= 3))
        neighbors = neighbors.allocate(init_state)
        init_fn, step_fn = self.simulator_init(efn, shift_fn, **self.simulator_params.sim_init_fn)
        state = init_fn(key, init_state, unbonded_neighbors=neighbors.idx, **self.simulator_params.init_fn)
        dev, dtype = init_state.center.device, init_state.center.dtype
        n = init_state.center.shape[0]
        traj = (torch.empty((n_steps, n, 3), device=dev, dtype=dtype), torch.empty((n_steps, n, 4), device=dev, dtype=dtype))
        # a list that never changes, or one that updates itself on the device in place (utils.neighbors: warp-slot layout
        # with the conditional rebuild): both are fixed launch sequences, so the step is captured in a CUDA graph
        static_list = isinstance(neighbors, NoNeighborList) or getattr(neighbors, "slots", None) is not None
        if self.use_cuda_graph and static_list and hasattr(step_fn, "launch"):
            self._run_graph(step_fn, state, neighbors, traj, n_steps)
        else:
            for _ in range(n_steps):
                state = step_fn(state, unbonded_neighbors=neighbors.idx, traj=traj)
                neighbors = neighbors.update(state.position.center)
        kT = self.simulator_params.kT
        return SimulatorTrajectory(
            center=traj[0], orientation=Quaternion(traj[1]), temperature=torch.full((n_steps,), float(kT), device=dev, dtype=dtype)
        )

    @staticmethod
    def _run_graph(step_fn, state, neighbors, traj, n_steps):
        """Steady-state step = [B(dt) A O A + row store + counter++] -> [energy + forces]; captured once, replayed."""
        f = step_fn.forces_for(state.position, neighbors.idx)
        c, q = state.position.center, state.position.orientation.vec

        def one_step():
            # 2 launches per step: B-A-O-A (zeroes the gradient buffers after the kick, bumps the step counter) and the pair
            # kernel (bonded + unbonded blocks in one grid, accumulating (dE/dcenter, dE/dquat) into the state's buffers)
            step_fn.launch(state, 2 if one_step.started else 0, traj=traj, advance=True, zero_forces=True)
            one_step.started = True
            if dynamic:  # third launch: displacement test + (rarely) the rebuild of the list, in place, on the device
                neighbors.update(c)
            f.accumulate_into(c, q, state.force.center, state.force.orientation.vec)

        one_step.started = False
        dynamic = getattr(neighbors, "slots", None) is not None
        stream = torch.cuda.Stream(device=c.device)
        stream.wait_stream(torch.cuda.current_stream(c.device))
        with torch.cuda.stream(stream):
            one_step()  # first step eagerly (phase 0: opening half kick only)
            if n_steps > 1:
                one_step()  # warm the allocator for the steady-state shape
            done = min(2, n_steps)
            # steady state: graphs of `block` steps (one host launch per block keeps the replay loop off the CPU's
            # critical path), then single-step replays for the remainder
            block = 64  # (fewer host launches per simulated time: the replay loop stays off a slow host's critical path)
            if n_steps - done >= 2 * block:
                big = torch.cuda.CUDAGraph()
                with torch.cuda.graph(big, stream=stream):
                    for _ in range(block):
                        one_step()
                for _ in range((n_steps - done) // block):
                    big.replay()
                done += ((n_steps - done) // block) * block
            if n_steps > done:
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=stream):
                    one_step()
                for _ in range(n_steps - done):
                    graph.replay()
            step_fn.launch(state, 1)  # closing half kick of the last step
        torch.cuda.current_stream(c.device).wait_stream(stream)
