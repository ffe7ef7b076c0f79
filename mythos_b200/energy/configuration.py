"""Configuration base class for energy terms (host side, unchanged contract).

Mirrors ``mythos/energy/configuration.py:17-123``: required / dependent / optimisable parameter bookkeeping,
``init_params`` (independent -> dependent, runs on every parameter update), ``|`` merge, ``from_dict``,
``to_dictionary`` and ``opt_params``.  Values are Python floats or 0-d / (4,4) float64 torch tensors; a tensor
with ``requires_grad`` flows through ``init_params`` into the packed parameter bank, which is how
d/dtheta is chained to the kernel's d/dparams (the reference does the same chain inside jax.grad).
"""

from __future__ import annotations

import copy
import warnings
from typing import Any

ERR_MISSING_REQUIRED_PARAMS = "Required properties {props} are not initialized."
ERR_OPT_DEPENDENT_PARAMS = "Only {req_params} permitted for optimization, but found {given_params}"
WARN_INIT_PARAMS_NOT_IMPLEMENTED = "init_params not implemented"
WARN_DEPENDENT_PARAMS_NOT_INITIALIZED = "Dependent parameters not initialized"

_META = ("params_to_optimize", "required_params", "non_optimizable_required_params", "dependent_params", "OPT_ALL")


class BaseConfiguration:
    """Dict-backed frozen record; subclasses declare ``required_params``, ``dependent_params``, ``optional_params``."""

    required_params: tuple[str, ...] = ()
    dependent_params: tuple[str, ...] = ()
    optional_params: tuple[str, ...] = ()
    non_optimizable_required_params: tuple[str, ...] = ()
    OPT_ALL: tuple[str, ...] = ("*",)
    term: str = ""  # name of the kernel-level term table these parameters pack into

    def __init__(self, params_to_optimize: tuple[str, ...] = (), **values: Any) -> None:
        object.__setattr__(self, "params_to_optimize", tuple(params_to_optimize))
        for meta in ("required_params", "dependent_params", "non_optimizable_required_params"):
            if meta in values:
                object.__setattr__(self, meta, tuple(values.pop(meta)))
        fields = self.field_names()
        unknown = set(values) - set(fields)
        if unknown:
            raise TypeError(f"{type(self).__name__} got unexpected parameters {sorted(unknown)}")
        object.__setattr__(self, "_values", {k: values.get(k) for k in fields})
        self.__post_init__()

    # -- record behaviour ------------------------------------------------------------------------------
    def field_names(self) -> tuple[str, ...]:
        seen: dict[str, None] = {}
        for k in (*self.required_params, *self.optional_params, *self.dependent_params):
            seen.setdefault(k, None)
        return tuple(seen)

    def __getattr__(self, name: str) -> Any:
        values = self.__dict__.get("_values")
        if values is not None and name in values:
            return values[name]
        raise AttributeError(name)

    def __setattr__(self, name: str, value: Any) -> None:
        raise AttributeError(f"{type(self).__name__} is frozen; use replace()")

    def __contains__(self, name: str) -> bool:
        return name in self._values or name in _META

    def keys(self):
        return [*self._values.keys(), *_META]

    def items(self):
        return [(k, getattr(self, k)) for k in self.keys()]

    def __iter__(self):
        return iter(self.keys())

    def replace(self, **changes: Any) -> "BaseConfiguration":
        new = copy.copy(self)
        object.__setattr__(new, "_values", dict(self._values))
        for k, v in changes.items():
            if k in _META:
                object.__setattr__(new, k, tuple(v))
            elif k in new._values:
                new._values[k] = v
            else:
                raise TypeError(f"{type(self).__name__} has no parameter {k!r}")
        new.__post_init__()
        return new

    def __getstate__(self):
        return self.__dict__

    def __setstate__(self, state):
        self.__dict__.update(state)

    def __repr__(self) -> str:
        inner = ", ".join(f"{k}={v!r}" for k, v in self._values.items() if v is not None)
        return f"{type(self).__name__}({inner})"

    # -- reference API -----------------------------------------------------------------------------------
    def __post_init__(self) -> None:
        missing = [p for p in self.required_params if self._values.get(p) is None]
        if missing:
            raise ValueError(ERR_MISSING_REQUIRED_PARAMS.format(props=",".join(missing)))
        optimizable = set(self.required_params) - set(self.non_optimizable_required_params)
        bad = set(self.params_to_optimize) - optimizable
        if bad and bad != set(self.OPT_ALL):
            raise ValueError(
                ERR_OPT_DEPENDENT_PARAMS.format(
                    req_params=",".join(sorted(optimizable)), given_params=",".join(sorted(bad))
                )
            )

    @property
    def opt_params(self) -> dict[str, Any]:
        if self.params_to_optimize == self.OPT_ALL:
            return {
                k: v
                for k, v in self._values.items()
                if k in self.required_params and k not in self.non_optimizable_required_params
            }
        return {k: v for k, v in self._values.items() if k in self.params_to_optimize}

    def init_params(self) -> "BaseConfiguration":
        warnings.warn(WARN_INIT_PARAMS_NOT_IMPLEMENTED, stacklevel=1)
        return self

    @classmethod
    def from_dict(cls, params: dict[str, Any], params_to_optimize: tuple[str, ...] = ()) -> "BaseConfiguration":
        return cls(**(dict(params) | {"params_to_optimize": params_to_optimize}))

    def to_dictionary(self, *, include_dependent: bool, exclude_non_optimizable: bool) -> dict[str, Any]:
        params = {k: getattr(self, k) for k in self.required_params}
        if include_dependent:
            for k in self.dependent_params:
                if (val := getattr(self, k)) is not None:
                    params[k] = val
                else:
                    warnings.warn(WARN_DEPENDENT_PARAMS_NOT_INITIALIZED, stacklevel=1)
        if exclude_non_optimizable:
            for k in self.non_optimizable_required_params:
                params.pop(k, None)
        return params

    def __or__(self, other: "BaseConfiguration | dict[str, Any]") -> "BaseConfiguration":
        if isinstance(other, BaseConfiguration):
            return self.replace(**{k: v for k, v in other._values.items() if v is not None and k in self._values})
        if isinstance(other, dict):
            return self.replace(**other)
        return NotImplemented
