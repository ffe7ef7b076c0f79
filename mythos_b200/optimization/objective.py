"""Objectives: Boltzmann reweighting of stored frames (DiffTRe), frame-sharded across the GPUs of one box.

Interface of ``mythos/optimization/objective.py:53-389`` -- ``ObjectiveOutput``, ``Objective``,
``compute_weights_and_neff``, ``compute_min_segment_neff``, ``compute_loss``, ``compute_loss_and_grad``,
``DiffTReObjective.calculate`` with the same state machine (needs_update / neff threshold / opt_steps).

What changed underneath: ``energy_fn.map(states)`` is one batched launch group whose forward also
writes the dE/dparams row of every frame, so ``value_and_grad(compute_loss)`` costs ONE pass over the frames instead of
the reference's forward + remat-forward + backward; and when ``torch.distributed`` is initialised the frames are
sharded contiguously over the ranks, per-frame energies are all-gathered (64 KiB at F=8192) so every rank sees the
full weight vector, and the weighted parameter gradients are summed with an all-reduce (NCCL over NVLink).
"""

from __future__ import annotations

import ctypes as C
import dataclasses as dc
import math
import typing
from collections.abc import Callable

import torch
import torch.distributed as dist

from mythos_b200 import _lib
from mythos_b200.energy import theta_tape
from mythos_b200.energy.base import EnergyFunction
from mythos_b200.rigid_body import Quaternion, RigidBody
from mythos_b200.simulators.io import SimulatorTrajectory

ERR_DIFFTRE_MISSING_KWARGS = "Missing required kwargs: {missing_kwargs}."
ERR_MISSING_ARG = "Missing required argument: {missing_arg}."
ERR_OBJECTIVE_NOT_READY = "Not all required observables have been obtained."


@dc.dataclass(frozen=True, kw_only=True)
class ObjectiveOutput:
    is_ready: bool
    grads: typing.Any = None
    observables: dict[str, typing.Any] = dc.field(default_factory=dict)
    state: dict[str, typing.Any] = dc.field(default_factory=dict)
    needs_update: tuple[str, ...] = dc.field(default_factory=tuple)


@dc.dataclass(frozen=True, kw_only=True)
class Objective:
    """Computes gradients from observables (``objective.py:53-136``)."""

    name: str
    required_observables: tuple[str, ...]
    logging_observables: tuple[str, ...] = dc.field(default_factory=tuple)
    grad_or_loss_fn: Callable = dc.field(repr=False, default=None)

    def __post_init__(self) -> None:
        if self.name is None:
            raise ValueError(ERR_MISSING_ARG.format(missing_arg="name"))
        if self.required_observables is None:
            raise ValueError(ERR_MISSING_ARG.format(missing_arg="required_observables"))
        if self.grad_or_loss_fn is None:
            raise ValueError(ERR_MISSING_ARG.format(missing_arg="grad_or_loss_fn"))

    def calculate(self, observables: dict[str, typing.Any], opt_params=None, **_kwargs) -> ObjectiveOutput:
        missing = [obs for obs in self.required_observables if obs not in observables]
        if missing:
            return ObjectiveOutput(is_ready=False, needs_update=tuple(missing))
        sorted_obs = [observables[key] for key in self.required_observables]
        grads, aux = self.grad_or_loss_fn(*sorted_obs)
        out = dict(aux)
        out.update(dict(zip(self.required_observables, sorted_obs, strict=True)))
        return ObjectiveOutput(is_ready=True, grads=grads, observables=out, state={}, needs_update=())

    def get_logging_observables(self, observables: dict[str, typing.Any]) -> list[tuple[str, typing.Any]]:
        return [(name, observables[name]) for name in self.logging_observables if name in observables]


# ------------------------------------------------------------------------------------------- reweighting
def _weights_kernel(beta: torch.Tensor, new_e: torch.Tensor, ref_e: torch.Tensor):
    """Forward of the weights through the CUDA kernel (mythos_b200_weights_neff_*): (weights (F,), n_eff)."""
    F = new_e.shape[0]
    dev, dtype = new_e.device, new_e.dtype
    beta = torch.as_tensor(beta, dtype=dtype, device=dev).expand(F).contiguous()
    w = torch.empty(F, dtype=dtype, device=dev)
    sums = torch.empty(4, dtype=dtype, device=dev)
    a = _lib.WeightsArgs()
    a.n_frames = F
    a.beta, a.e_new, a.e_ref = beta.data_ptr(), new_e.contiguous().data_ptr(), ref_e.to(dtype).contiguous().data_ptr()
    a.weights, a.sums = w.data_ptr(), sums.data_ptr()
    fn = getattr(_lib.lib(), f"mythos_b200_weights_neff_{_lib.suffix(dtype)}")
    with torch.cuda.device(dev):
        _lib.check(fn(_lib.current_stream(dev), C.byref(a)), "mythos_b200_weights_neff")
    return w, sums[3]


class _Weights(torch.autograd.Function):
    """w = softmax(-beta (E - Eref)); n_eff = exp(-sum w ln w)/F, with the analytic softmax / entropy backward."""

    @staticmethod
    def forward(ctx, beta, new_e, ref_e):
        w, neff = _weights_kernel(beta, new_e.detach(), ref_e.detach())
        b = torch.as_tensor(beta, dtype=new_e.dtype, device=new_e.device).expand_as(new_e)
        ctx.save_for_backward(w, neff, b)
        ctx.set_materialize_grads(False)  # an unused output arrives as None, not as a zero tensor
        return w, neff

    @staticmethod
    def backward(ctx, g_w, g_neff):
        w, neff, beta = ctx.saved_tensors
        # x_k = -beta_k (E_k - Eref_k);  dw_j/dx_k = w_j (delta_jk - w_k);  S = -sum w ln w, dS/dx_k = -w_k (ln w_k + S)
        gx = torch.zeros_like(w) if g_w is None else w * (g_w - (g_w * w).sum())
        if g_neff is not None:
            wlw = torch.xlogy(w, w)  # 0 * log 0 = 0: a weight that underflowed must not poison the gradient with NaN
            S = -wlw.sum()
            gx = gx + g_neff * neff * (-(wlw + w * S))
        g_e = -beta * gx
        return None, g_e, -g_e


def compute_weights_and_neff(beta, new_energies: torch.Tensor, ref_energies: torch.Tensor):
    """Weights and normalised effective sample size of a trajectory (``objective.py:139-163``, DiffTRe eqs 4-5)."""
    _lib.require_cuda(new_energies, "new_energies")
    return _Weights.apply(beta, new_energies, ref_energies)


def compute_min_segment_neff(temperature: torch.Tensor, new_energies: torch.Tensor, ref_energies: torch.Tensor) -> float:
    """Minimum n_eff over temperature segments (``objective.py:166-195``)."""

    def segment_neff(temp) -> float:
        mask = temperature == temp
        _, neff = compute_weights_and_neff(1.0 / temp, new_energies[mask].detach(), ref_energies[mask].detach())
        return float(neff)

    return min(segment_neff(t) for t in torch.unique(temperature))


# ------------------------------------------------------------------------------------------- frame sharding
def _world() -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def all_ranks_ok(good: bool) -> bool:
    """Logical AND of a per-rank flag over the ranks (a repeated pass re-enters the collectives, so the ranks must agree)."""
    rank, world = _world()
    if world == 1:
        return good
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([1 if good else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t.item())


def shard_bounds(n_frames: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block of frames owned by ``rank`` (remainder spread over the first ranks)."""
    base, rem = divmod(n_frames, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_frames(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """all_gather of contiguous per-rank blocks (sizes may differ by one: blocks are padded to the widest)."""
    rank, world = _world()
    sizes = [hi - lo for lo, hi in (shard_bounds(n_total, r, world) for r in range(world))]
    width = max(sizes)
    if min(sizes) == width and local.is_contiguous():  # equal blocks (the usual case): one collective straight into the result
        out = torch.empty(width * world, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local)
        return out
    padded = torch.zeros(width, dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    chunks = [torch.empty(width, dtype=local.dtype, device=local.device) for _ in range(world)]
    dist.all_gather(chunks, padded)
    return torch.cat([c[:s] for c, s in zip(chunks, sizes)])


class _GatherFrames(torch.autograd.Function):
    """all_gather of per-rank frame energies into the full (F,) vector; backward keeps this rank's slice of the
    cotangent (every rank evaluates the same loss on the same full vector, so no reduction is needed here)."""

    @staticmethod
    def forward(ctx, local, n_total):
        rank, world = _world()
        ctx.bounds = shard_bounds(n_total, rank, world)
        return gather_frames(local, n_total)

    @staticmethod
    def backward(ctx, g):
        lo, hi = ctx.bounds
        return g[lo:hi].contiguous(), None


def sharded_map(energy_fn: EnergyFunction, states: RigidBody, sharded: bool | None = None, observables=None) -> torch.Tensor:
    """``energy_fn.map(states)`` -> (F,) with the frames split over the ranks when torch.distributed is up.

    ``states`` holds ALL frames on every rank (they are a few hundred MB and arrive by the same route on each rank);
    each rank evaluates its contiguous block and the blocks are all-gathered."""
    rank, world = _world()
    if sharded is None:
        sharded = world > 1
    if not sharded or world == 1:
        return energy_fn.map(states, observables=observables) if observables is not None else energy_fn.map(states)
    shard = getattr(states, "shard", None)
    if shard is not None:  # the caller already holds only this rank's block (lo, hi, total)
        lo, hi, total = shard
        if (lo, hi) != shard_bounds(total, rank, world):
            raise ValueError(f"states.shard {shard} is not rank {rank}'s block of {total} frames over {world} ranks")
        return _GatherFrames.apply(energy_fn.map(states), total)
    lo, hi = shard_bounds(states.center.shape[0], rank, world)
    local = energy_fn.map(RigidBody(states.center[lo:hi], Quaternion(states.orientation.vec[lo:hi])))
    return _GatherFrames.apply(local, states.center.shape[0])


def allreduce_grads(grads: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]:
    """Sum the per-rank parameter gradients (each rank holds sum over ITS frames) with one all-reduce."""
    rank, world = _world()
    if world == 1 or not grads:
        return grads
    keys = sorted(grads)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    # parameters may be scalars or tables (ss_stack_weights (4,4), ...): flattened into one buffer, one collective
    sizes = [grads[k].numel() for k in keys]
    flat = torch.cat([grads[k].reshape(-1).to(torch.float64) for k in keys]).to(dev)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    back: dict = {}  # one copy per destination device, not one per parameter
    out = {}
    off = 0
    for k, sz in zip(keys, sizes):
        d = grads[k].device
        if d not in back:
            back[d] = flat.to(d)
        out[k] = back[d][off:off + sz].reshape(grads[k].shape).to(grads[k].dtype)
        off += sz
    return out


# ------------------------------------------------------------------------------------------- loss
def compute_loss(opt_params, energy_fn: EnergyFunction, beta, loss_fn, ref_states, ref_energies, observables):
    """loss, (n_eff, measured value, new energies) (``objective.py:198-232``)."""
    if (isinstance(getattr(ref_states, "center", None), torch.Tensor) and not ref_states.center.is_cuda
            and (_world()[1] == 1 or getattr(ref_states, "shard", None) is not None)):
        # pinned host frames: their first chunk crosses PCIe while the host evaluates theta -> parameter bank below
        from mythos_b200.energy import functional

        functional.prefetch_frames(ref_states.center, ref_states.orientation.vec)
    # = energy_fn.with_params(opt_params); the theta -> parameter-bank chain is replayed from a tape recorded once per
    # energy function (what jit does for the reference) instead of ~700 eager autograd nodes per step
    base_fn = energy_fn
    energy_fn = theta_tape.bind(energy_fn, opt_params)
    # observables the loss function declares (loss_fn.fused_observables = ObservableSet([...])) are evaluated by the epilogue of
    # the same kernel that computes the energies; loss_fn's own observable(ref_states) calls then find them ready
    new_energies = sharded_map(energy_fn, ref_states, observables=getattr(loss_fn, "fused_observables", None))
    weights, neff = compute_weights_and_neff(beta, new_energies, ref_energies)
    loss_params, loss_efn = opt_params, energy_fn
    world = _world()[1]
    if world > 1 and torch.is_grad_enabled():
        # Frame sharding: the gradient that flows through the frames is this rank's share (the gather's backward keeps the
        # rank's slice) and the shares are SUMMED over the ranks.  Whatever loss_fn takes from theta directly -- through
        # opt_params or through the energy function it is handed -- is identical on every rank, so it is given to loss_fn
        # with its gradient scaled by 1 / world: the same sum then returns it exactly once, without a second, frames-free
        # pass over loss_fn.
        loss_params = _scale_param_grads(opt_params, 1.0 / world)
        loss_efn = _LazyWithParams(base_fn, loss_params)
    loss, (measured_value, _) = loss_fn(ref_states, weights, loss_efn, loss_params, observables)
    return loss, (neff, measured_value, new_energies)


def _scale_grad(t: torch.Tensor, s: float) -> torch.Tensor:
    """Same value as ``t``, gradient scaled by ``s``."""
    if not (isinstance(t, torch.Tensor) and t.requires_grad):
        return t
    d = t.detach()
    return d + (t - d) * s


def _scale_param_grads(opt_params, s: float):
    if isinstance(opt_params, theta_tape.FlatParams):
        return opt_params.like(_scale_grad(opt_params.flat, s))
    return {k: _scale_grad(v, s) for k, v in opt_params.items()}


def compute_loss_and_grad(opt_params: dict[str, torch.Tensor], energy_fn, beta, loss_fn, ref_states, ref_energies, observables):
    """``jax.value_and_grad(compute_loss, has_aux=True)`` (``objective.py:235``) for a dict of float tensors (scalars or
    tables).  The parameters are differentiated as ONE flat leaf (``theta_tape.FlatParams``): ``loss_fn`` still sees a
    mapping name -> tensor, the autograd engine sees one leaf instead of a hundred."""
    from mythos_b200.energy import functional

    if not all(isinstance(v, (int, float, torch.Tensor)) for v in opt_params.values()):
        # parameter PYTREES (e.g. pseq = (unpaired (n_u,4), base-paired (n_bp,4)) for sequence design): one leaf per tensor
        return _loss_and_grad_pytree(opt_params, energy_fn, beta, loss_fn, ref_states, ref_energies, observables)
    for attempt in range(functional.MAX_PASS_REPEATS + 1):
        if attempt == functional.MAX_PASS_REPEATS:
            raise _lib.MythosB200Error(f"pair lists still overflow after {attempt} passes over the reference states")
        leaves = theta_tape.FlatParams(opt_params)
        leaves.flat.requires_grad_(True)
        # the pair-list overflow flags are read after the backward (which syncs anyway), not in the middle of the pass
        with functional.deferred_verification() as checks:
            loss, aux = compute_loss(leaves, energy_fn, beta, loss_fn, ref_states, ref_energies, observables)
            checks.enqueue()  # the flags' read-back is in flight before the backward's own device -> host copy waits
            (g,) = torch.autograd.grad(loss, [leaves.flat], allow_unused=True)
        good = checks.ok()
        g = torch.zeros_like(leaves.flat) if g is None else g
        if _world()[1] > 1:
            # each rank holds (frames part of ITS block) + (direct part / world), see compute_loss: one SUM all-reduce.
            # The ranks' "my lists overflowed" flags ride in the same collective (no collective + host read of their own).
            red = allreduce_grads({"flat": g, "overflowed": torch.tensor([0.0 if good else 1.0], dtype=torch.float64)})
            g, good = red["flat"], float(red["overflowed"][0]) == 0.0
        if good:
            break
    grads = leaves.unflatten(g)
    for k, v in grads.items():  # (the flat leaf is float64; only parameters given in another floating type are converted)
        p = opt_params[k]
        if isinstance(p, torch.Tensor) and p.is_floating_point() and p.dtype != v.dtype:
            grads[k] = v.to(p.dtype)
    return (loss.detach(), tuple(a.detach() if isinstance(a, torch.Tensor) else a for a in aux)), grads


def _loss_and_grad_pytree(opt_params, energy_fn, beta, loss_fn, ref_states, ref_energies, observables):
    """``compute_loss_and_grad`` for parameter values that are nested containers of tensors (single process)."""
    from torch.utils import _pytree as pytree

    from mythos_b200.energy import functional

    if _world()[1] > 1:
        raise NotImplementedError("pytree-valued parameters (pseq) with frame sharding")
    flat, spec = pytree.tree_flatten(opt_params)
    for attempt in range(functional.MAX_PASS_REPEATS + 1):
        if attempt == functional.MAX_PASS_REPEATS:
            raise _lib.MythosB200Error(f"pair lists still overflow after {attempt} passes over the reference states")
        leaves = [torch.as_tensor(v, dtype=torch.float64).detach().clone().requires_grad_(True) for v in flat]
        with functional.deferred_verification() as checks:
            loss, aux = compute_loss(pytree.tree_unflatten(leaves, spec), energy_fn, beta, loss_fn, ref_states, ref_energies, observables)
            checks.enqueue()
            gl = torch.autograd.grad(loss, leaves, allow_unused=True)
        if all_ranks_ok(checks.ok()):
            break
    grads = pytree.tree_unflatten([torch.zeros_like(x) if g is None else g for x, g in zip(leaves, gl)], spec)
    return (loss.detach(), tuple(a.detach() if isinstance(a, torch.Tensor) else a for a in aux)), grads


class _LazyWithParams:
    """``energy_fn.with_params(params)`` evaluated on first use: the frames-free pass below hands it to ``loss_fn``, and
    most loss functions never touch it -- the theta -> parameter chain (milliseconds of host time) is then not rebuilt."""

    def __init__(self, energy_fn, params):
        object.__setattr__(self, "_make", lambda: energy_fn.with_params(params))
        object.__setattr__(self, "_built", None)

    def _get(self):
        if self._built is None:
            object.__setattr__(self, "_built", self._make())
        return self._built

    def __getattr__(self, name):
        return getattr(self._get(), name)

    def __call__(self, *args, **kwargs):
        return self._get()(*args, **kwargs)


_REFERENCE_STATES: dict = {}  # (ids of the trajectories' frame tensors, n_equilibration) -> (weak refs, sliced + concatenated states)


def _reference_states(trajectories: list, n_equilibration_steps: int) -> SimulatorTrajectory:
    """Equilibration slicing + concatenation of ``calculate`` (``objective.py:322-326``), remembered per set of input
    trajectories: an optimiser calls ``calculate`` many times on the SAME stored trajectories before it resamples, and
    handing the same tensor objects down each time is what lets the evaluation keep its per-frame pair lists
    (``functional._PairListCache``) instead of rebuilding them on every call."""
    import weakref

    key = (tuple(id(t.center) for t in trajectories), int(n_equilibration_steps))
    hit = _REFERENCE_STATES.get(key)
    if hit is not None and all(r() is t.center and v == t.center._version for (r, v), t in zip(hit[0], trajectories)):
        return hit[1]
    if n_equilibration_steps > 0:
        sliced = [obs.slice(slice(n_equilibration_steps, obs.length(), None)) for obs in trajectories]
    else:
        sliced = list(trajectories)
    states = SimulatorTrajectory.concat(sliced)
    for k in [k for k, (refs, _) in _REFERENCE_STATES.items() if any(r() is None for r, _ in refs)]:
        del _REFERENCE_STATES[k]
    if len(_REFERENCE_STATES) > 8:
        _REFERENCE_STATES.pop(next(iter(_REFERENCE_STATES)))
    _REFERENCE_STATES[key] = ([(weakref.ref(t.center), t.center._version) for t in trajectories], states)
    return states


@dc.dataclass(frozen=True, kw_only=True)
class DiffTReObjective(Objective):
    """DiffTRe gradient computation with the reference's state machine (``objective.py:239-389``)."""

    energy_fn: EnergyFunction = dc.field(repr=False, default=None)
    n_equilibration_steps: int = 0
    min_n_eff_factor: float = 0.95
    max_valid_opt_steps: float = math.inf

    def __post_init__(self) -> None:
        Objective.__post_init__(self)
        if self.energy_fn is None:
            raise ValueError(ERR_MISSING_ARG.format(missing_arg="energy_fn"))
        if self.n_equilibration_steps is None:
            raise ValueError(ERR_MISSING_ARG.format(missing_arg="n_equilibration_steps"))
        if self.n_equilibration_steps < 0:
            raise ValueError(f"n_equilibration_steps must be non-negative, got {self.n_equilibration_steps}.")
        if self.max_valid_opt_steps <= 0:
            raise ValueError("max_valid_opt_steps must be positive or infinity.")

    def calculate(self, observables, opt_params, opt_steps: int = 0, reference_opt_params=None) -> ObjectiveOutput:
        if opt_steps >= self.max_valid_opt_steps:
            return ObjectiveOutput(is_ready=False, needs_update=tuple(self.required_observables), state={"opt_steps": 0})
        missing = [obs for obs in self.required_observables if obs not in observables]
        if missing:
            return ObjectiveOutput(is_ready=False, needs_update=tuple(missing))
        sorted_obs = [observables[key] for key in self.required_observables]
        trajectories = [obs for obs in sorted_obs if isinstance(obs, SimulatorTrajectory)]
        if not trajectories:
            raise ValueError("No SimulatorTrajectory observables found in observables.")
        reference_states = _reference_states(trajectories, self.n_equilibration_steps)
        if reference_states.length() == 0:
            raise ValueError("Equilibration slicing yields no states! Note slicing is in number of snapshots, not timesteps.")
        if reference_states.temperature is None:
            raise ValueError(
                "SimulatorTrajectory.temperature is None. DiffTRe requires per-state temperature (kT) on the trajectory."
            )
        beta = 1.0 / reference_states.temperature
        reference_opt_params = reference_opt_params or opt_params
        with torch.no_grad():
            reference_energies = sharded_map(theta_tape.bind(self.energy_fn, reference_opt_params), reference_states)
            same = reference_opt_params is opt_params
            new_energies = reference_energies if same else sharded_map(theta_tape.bind(self.energy_fn, opt_params), reference_states)
        neff = compute_min_segment_neff(reference_states.temperature, new_energies, reference_energies)
        if neff < self.min_n_eff_factor:
            return ObjectiveOutput(
                is_ready=False, needs_update=tuple(self.required_observables), observables={"neff": neff}, state={"opt_steps": 0}
            )
        (loss, (_, measured_value, _)), grads = compute_loss_and_grad(
            opt_params, self.energy_fn, beta, self.grad_or_loss_fn, reference_states, reference_energies, sorted_obs
        )
        return ObjectiveOutput(
            is_ready=True,
            grads=grads,
            observables={"loss": loss, "neff": neff, measured_value[0]: measured_value[1]},
            state={"opt_steps": opt_steps + 1, "reference_opt_params": reference_opt_params},
        )
