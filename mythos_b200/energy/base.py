"""Energy-function base classes: the reference's public interface over the fused CUDA kernels.

Interface kept from ``mythos/energy/base.py:25-462``: ``EnergyFunction`` (``__call__``, ``map``, ``with_params``,
``with_props``, ``with_noopt``, ``params_dict``, ``opt_params``), ``BaseEnergyFunction`` (fields ``params``,
``displacement_fn``, ``seq``, ``bonded_neighbors``, ``unbonded_neighbors``, ``topology``, ``transform_fn``; ``+`` and
``*`` composition), ``ComposedEnergyFunction`` (``compute_terms``, weights, ``without_terms``, ``from_lists``,
global parameter namespace) and ``QualifiedComposedEnergyFunction``.

What changed underneath: a term class no longer carries arithmetic.  It names a kernel term
(``TERM``), its functional form, and a configuration; ``__call__`` packs the configuration(s) into the
kernel-level parameter bank and launches the fused pair kernels once for *all* terms of a composition
(the reference traces T separate sub-graphs that each re-derive the sites, ``base.py:312-314``).
``map`` evaluates all frames in one batched launch instead of ``lax.map`` over chunks of 100.
Inputs are torch CUDA tensors; gradients w.r.t. positions, orientations and theta come from the analytic backward
kernels through torch.autograd (``mythos_b200.energy.functional``).
"""

from __future__ import annotations

import dataclasses as dc
from abc import ABC, abstractmethod
from collections.abc import Callable
from typing import Any, ClassVar, Union

import torch

from mythos_b200 import _lib
from mythos_b200.energy import functional, model as kmodel
from mythos_b200.energy.configuration import BaseConfiguration
from mythos_b200.rigid_body import Quaternion, RigidBody

ERR_PARAM_NOT_FOUND = "Parameter '{key}' not found in {class_name}"
ERR_CALL_NOT_IMPLEMENTED = "Subclasses must implement this method"
ERR_COMPOSED_ENERGY_FN_LEN_MISMATCH = "Weights must have the same length as energy functions"
ERR_COMPOSED_ENERGY_FN_TYPE_ENERGY_FNS = "energy_fns must be a list of energy functions"


class _Record:
    """dict-like view + replace() shared by the frozen dataclasses below (what chex.dataclass provides)."""

    def keys(self):
        return [f.name for f in dc.fields(self)]

    def __getitem__(self, key: str) -> Any:
        return getattr(self, key)

    def replace(self, **changes: Any):
        return dc.replace(self, **changes)


@dc.dataclass(frozen=True, kw_only=True)
class EnergyFunction(_Record, ABC):
    """Callable that maps a RigidBody to the scalar energy of the system.

    ``map_batch_size`` / ``map_checkpoint`` are accepted for interface compatibility; ``map`` always runs every
    frame in one batched launch and the backward always recomputes (the kernels save nothing).
    """

    map_batch_size: int | None = 100
    map_checkpoint: bool = True

    @abstractmethod
    def __call__(self, body: RigidBody) -> torch.Tensor: ...

    @abstractmethod
    def with_params(self, *repl_dicts: dict, **repl_kwargs: Any) -> "EnergyFunction": ...

    @abstractmethod
    def with_props(self, **kwargs) -> "EnergyFunction": ...

    @abstractmethod
    def with_noopt(self, *params: str) -> "EnergyFunction": ...

    @abstractmethod
    def params_dict(self, *, include_dependent: bool = True, exclude_non_optimizable: bool = False) -> dict: ...

    @abstractmethod
    def opt_params(self) -> dict[str, Any]: ...

    @abstractmethod
    def map(self, body_sequence: RigidBody) -> torch.Tensor:
        """Energy of every frame of a stacked RigidBody ``(F,N,3) / (F,N,4)`` -> ``(F,)``."""


def _frames(body: RigidBody) -> tuple[torch.Tensor, torch.Tensor, bool]:
    c, q = body.center, body.orientation.vec if isinstance(body.orientation, Quaternion) else body.orientation
    single = c.dim() == 2
    if single:
        c, q = c.unsqueeze(0), q.unsqueeze(0)
    return c, q, single


@dc.dataclass(frozen=True, kw_only=True)
class BaseEnergyFunction(EnergyFunction):
    """One energy term.  Subclasses set ``TERM`` (kernel term id) and ``FORM`` (variant flags)."""

    params: BaseConfiguration
    displacement_fn: Callable
    seq: Any = None
    bonded_neighbors: Any = None
    unbonded_neighbors: Any = None
    topology: dc.InitVar[Any] = None
    transform_fn: Callable | None = None

    TERM: ClassVar[int] = -1
    FORM: ClassVar[dict] = {}
    HYBRID: ClassVar[bool] = False  # NA1 classes: three parameter banks selected by nt_type

    def __post_init__(self, topology) -> None:
        if topology is not None:
            object.__setattr__(self, "seq", topology.seq)
            object.__setattr__(self, "bonded_neighbors", topology.bonded_neighbors)
            ub_t = getattr(topology, "unbonded_neighbors_t", None)
            object.__setattr__(self, "unbonded_neighbors", topology.unbonded_neighbors.T if ub_t is None else ub_t)
        elif any(x is None for x in (self.seq, self.bonded_neighbors, self.unbonded_neighbors)):
            raise ValueError("Missing topology information")

    @classmethod
    def create_from(cls, other: "EnergyFunction", **kwargs) -> "EnergyFunction":
        props = {k: other[k] for k in other.keys() if k in {f.name for f in dc.fields(cls)}} | kwargs
        return cls(**props)

    def __add__(self, other: "BaseEnergyFunction") -> "ComposedEnergyFunction":
        if not isinstance(other, BaseEnergyFunction):
            return NotImplemented
        return ComposedEnergyFunction(energy_fns=[self, other])

    def __mul__(self, other: float) -> "ComposedEnergyFunction":
        if not isinstance(other, float | int):
            return NotImplemented
        return ComposedEnergyFunction(energy_fns=[self], weights=torch.tensor([float(other)], dtype=torch.float64))

    def with_props(self, **kwargs: Any) -> "EnergyFunction":
        return self.replace(**kwargs)

    def with_noopt(self, *params: str) -> "EnergyFunction":
        updated = set(self.params.non_optimizable_required_params) | set(params)
        return self.replace(params=self.params.replace(non_optimizable_required_params=list(updated)))

    def opt_params(self) -> dict[str, Any]:
        return self.params.opt_params

    def with_params(self, *repl_dicts: dict, **repl_kwargs: Any) -> "EnergyFunction":
        new_params = self.params
        for replacements in repl_dicts:
            new_params = new_params | replacements
        new_params = new_params | repl_kwargs
        return self.replace(params=new_params.init_params())

    def params_dict(self, include_dependent: bool = True, exclude_non_optimizable: bool = False) -> dict:
        return self.params.to_dictionary(
            include_dependent=include_dependent, exclude_non_optimizable=exclude_non_optimizable
        )

    # -- evaluation --------------------------------------------------------------------------------------
    def extra_topology(self) -> dict:
        """Per-nucleotide integer arrays beyond seq / bonds that this term needs (is_end, nt_type)."""
        out = {}
        nt = getattr(self.params, "nt_type", None) if "nt_type" in self.params else None
        if nt is not None:
            out["nt_type"] = nt
        return out

    def compute_energy_frames(self, body: RigidBody) -> torch.Tensor:
        """(F,) energies of this single term."""
        c, q, _ = _frames(body)
        plan = kmodel.plan_for([self])
        terms = plan.evaluate(c, q)
        return terms[:, self.TERM]

    def __call__(self, body: RigidBody) -> torch.Tensor:
        e = self.compute_energy_frames(body)
        return e[0] if body.center.dim() == 2 else e

    def compute_energy(self, nucleotide: RigidBody) -> torch.Tensor:
        """Kept for interface parity (``base.py:210-212``): the energy of an (already given) rigid body."""
        return self(nucleotide)

    def map(self, body_sequence: RigidBody) -> torch.Tensor:
        return self.compute_energy_frames(body_sequence)


@dc.dataclass(frozen=True)
class ComposedEnergyFunction(EnergyFunction):
    """Linear combination of energy terms sharing one global parameter namespace (``base.py:216-434``)."""

    energy_fns: list[BaseEnergyFunction] = dc.field(default_factory=list)
    weights: torch.Tensor | None = None
    strict_params: bool = True

    def __post_init__(self) -> None:
        if not isinstance(self.energy_fns, list) or not all(isinstance(fn, BaseEnergyFunction) for fn in self.energy_fns):
            raise TypeError(ERR_COMPOSED_ENERGY_FN_TYPE_ENERGY_FNS)
        if self.weights is not None and len(self.weights) != len(self.energy_fns):
            raise ValueError(ERR_COMPOSED_ENERGY_FN_LEN_MISMATCH)

    def with_props(self, **kwargs: Any) -> "ComposedEnergyFunction":
        return self.replace(energy_fns=[fn.with_props(**kwargs) for fn in self.energy_fns])

    def _param_in_fn(self, param: str, fn: BaseEnergyFunction) -> bool:
        return param in fn.params

    def _rename_param_for_fn(self, param: str, _fn: BaseEnergyFunction) -> str:
        return param

    def _rename_param_from_fn(self, param: str, _fn: BaseEnergyFunction) -> str:
        return param

    def with_noopt(self, *params: str) -> "ComposedEnergyFunction":
        energy_fns = []
        for fn in self.energy_fns:
            fn_params = [self._rename_param_for_fn(p, fn) for p in params if self._param_in_fn(p, fn)]
            energy_fns.append(fn.with_noopt(*fn_params))
        return self.replace(energy_fns=energy_fns)

    def opt_params(self, from_fns: list[type] | None = None) -> dict[str, Any]:
        fns = self.energy_fns if from_fns is None else [fn for fn in self.energy_fns if type(fn) in from_fns]
        return {self._rename_param_from_fn(k, fn): v for fn in fns for k, v in fn.opt_params().items()}

    def with_params(self, *repl_dicts: dict, **repl_kwargs: Any) -> "ComposedEnergyFunction":
        all_replacements = set(repl_kwargs) | {k for arg in repl_dicts for k in arg}
        used = set()
        energy_fns = []
        for fn in self.energy_fns:
            new_params = {k: v for arg in repl_dicts for k, v in arg.items() if self._param_in_fn(k, fn)}
            new_params.update({k: v for k, v in repl_kwargs.items() if self._param_in_fn(k, fn)})
            used.update(new_params.keys())
            new_params = {self._rename_param_for_fn(k, fn): v for k, v in new_params.items()}
            energy_fns.append(fn.with_params(**new_params))
        if self.strict_params and (unused := all_replacements - used):
            raise ValueError(f"Some parameters were not used in any energy function: {unused}.")
        return self.replace(energy_fns=energy_fns)

    def params_dict(self, *, include_dependent: bool = True, exclude_non_optimizable: bool = False) -> dict:
        params = {}
        for fn in self.energy_fns:
            fn_params = fn.params_dict(include_dependent=include_dependent, exclude_non_optimizable=exclude_non_optimizable)
            params.update({self._rename_param_from_fn(k, fn): v for k, v in fn_params.items()})
        return params

    # -- evaluation --------------------------------------------------------------------------------------
    def compute_terms_frames(self, body: RigidBody) -> torch.Tensor:
        """(F, T) per-term energies, T = len(energy_fns), from as few fused launches as the composition allows."""
        c, q, _ = _frames(body)
        cols: list[torch.Tensor | None] = [None] * len(self.energy_fns)
        for group in kmodel.fusable_groups(self.energy_fns):
            plan = kmodel.plan_for([self.energy_fns[k] for k in group])
            terms = plan.evaluate(c, q)
            for k in group:
                cols[k] = terms[:, self.energy_fns[k].TERM]
        return torch.stack(cols, dim=1)

    def compute_terms(self, body: RigidBody) -> torch.Tensor:
        t = self.compute_terms_frames(body)
        return t[0] if body.center.dim() == 2 else t

    def _combine(self, terms: torch.Tensor) -> torch.Tensor:
        if self.weights is None:
            return terms.sum(dim=-1)
        w = torch.as_tensor(self.weights, dtype=terms.dtype, device=terms.device)
        return terms @ w

    def __call__(self, body: RigidBody) -> torch.Tensor:
        return self._combine(self.compute_terms(body))

    def map(self, body_sequence: RigidBody, observables=None) -> torch.Tensor:
        """(F,) energies.  When the whole composition fuses into one launch group the weighted sum and the
        parameter-gradient rows come out of a single pass (``functional._FrameEnergy``, the DiffTRe shape).

        ``observables``: a ``mythos_b200.observables.ObservableSet`` evaluated in the SAME pass (epilogue of the
        frame-resident kernel); its members then answer ``observable(body_sequence)`` from that result."""
        c, q, _ = _frames(body_sequence)
        groups = kmodel.fusable_groups(self.energy_fns)
        if len(groups) == 1:
            w = torch.zeros(_lib.N_TERMS, dtype=torch.float64)
            wt = torch.ones(len(self.energy_fns), dtype=torch.float64) if self.weights is None else torch.as_tensor(self.weights, dtype=torch.float64).cpu()
            for k, fn in enumerate(self.energy_fns):
                w[fn.TERM] = wt[k]
            return kmodel.evaluate_with_observables(kmodel.plan_for(self.energy_fns), c, q, w, observables)
        return self._combine(self.compute_terms_frames(body_sequence))

    def without_terms(self, *terms: list[str | type]) -> "ComposedEnergyFunction":
        new_fns, new_w = [], []
        for i, fn in enumerate(self.energy_fns):
            if type(fn) in terms or fn.__class__.__name__ in terms:
                continue
            new_fns.append(fn)
            if self.weights is not None:
                new_w.append(self.weights[i])
        weights = None if self.weights is None else torch.as_tensor([float(w) for w in new_w], dtype=torch.float64)
        return self.replace(energy_fns=new_fns, weights=weights)

    def add_energy_fn(self, energy_fn: BaseEnergyFunction, weight: float = 1.0) -> "ComposedEnergyFunction":
        if self.weights is None:
            weights = None if weight == 1.0 else torch.tensor([1.0] * len(self.energy_fns) + [weight], dtype=torch.float64)
        else:
            weights = torch.cat([torch.as_tensor(self.weights, dtype=torch.float64), torch.tensor([weight], dtype=torch.float64)])
        return ComposedEnergyFunction(energy_fns=[*self.energy_fns, energy_fn], weights=weights)

    def add_composable_energy_fn(self, energy_fn: "ComposedEnergyFunction") -> "ComposedEnergyFunction":
        other_w = energy_fn.weights
        w_none, ow_none = self.weights is None, other_w is None
        if w_none and ow_none:
            weights = None
        else:
            mine = torch.ones(len(self.energy_fns), dtype=torch.float64) if w_none else torch.as_tensor(self.weights, dtype=torch.float64)
            theirs = torch.ones(len(energy_fn.energy_fns), dtype=torch.float64) if ow_none else torch.as_tensor(other_w, dtype=torch.float64)
            weights = torch.cat([mine, theirs])
        return ComposedEnergyFunction(energy_fns=self.energy_fns + energy_fn.energy_fns, weights=weights)

    def __add__(self, other: Union[BaseEnergyFunction, "ComposedEnergyFunction"]) -> "ComposedEnergyFunction":
        if isinstance(other, BaseEnergyFunction):
            return self.add_energy_fn(other)
        if isinstance(other, ComposedEnergyFunction):
            return self.add_composable_energy_fn(other)
        return NotImplemented

    def __radd__(self, other):
        return self.__add__(other)

    @classmethod
    def from_lists(
        cls,
        energy_fns: list[type[BaseEnergyFunction]],
        energy_configs: list[BaseConfiguration],
        weights: list[float] | None = None,
        **kwargs,
    ) -> "ComposedEnergyFunction":
        w = torch.ones(len(energy_fns), dtype=torch.float64) if weights is None else torch.as_tensor(weights, dtype=torch.float64)
        fns = [ef(**kwargs, params=ec.init_params()) for ef, ec in zip(energy_fns, energy_configs, strict=True)]
        return cls(energy_fns=fns, weights=w)


class QualifiedComposedEnergyFunction(ComposedEnergyFunction):
    """Parameters are addressed as ``ClassName.param`` instead of sharing one namespace (``base.py:437-462``)."""

    def _param_in_fn(self, param: str, fn: BaseEnergyFunction) -> bool:
        cls, param = param.split(".", 1)
        return param in fn.params and fn.__class__.__qualname__ == cls

    def _rename_param_for_fn(self, param: str, fn: BaseEnergyFunction) -> str:
        return param.split(".", 1)[1]

    def _rename_param_from_fn(self, param: str, fn: BaseEnergyFunction) -> str:
        return f"{fn.__class__.__qualname__}.{param}"
