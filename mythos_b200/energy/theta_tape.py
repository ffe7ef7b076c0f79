"""theta -> parameter bank as a replayable scalar tape (host side of every DiffTRe step).

The reference jits ``energy_fn.with_params(opt_params)`` together with the rest of the step
(``mythos/optimization/objective.py:224``), so the ``init_params`` chain of every term
(``mythos/energy/configuration.py:110-113``, ``dna1/base_smoothing_functions.py:48-142``) is compiled once and costs
nothing per step.  Run eagerly under torch autograd the same chain is ~700 tiny CPU ops forward and as many backward:
7.5 ms of every pass, more than the GPU needs for 1000 frames.  This module gives the eager host the same deal XLA gets:

* the chain of ONE energy function object is traced once (``make_fx`` over the unchanged Python ``init_params`` /
  ``pack_bank`` code), lowered to a straight-line tape of scalar float64 operations (common sub-expressions merged,
  constants folded, dead nodes dropped) and checked against the eager chain -- values and vector-Jacobian product;
* every later step replays the tape in C++ (``mythos_b200_theta_tape_forward`` / ``_vjp``, ``csrc/theta_tape.cu``):
  microseconds;
* ``bind(energy_fn, opt_params)`` returns an object that evaluates ``map`` with the replayed bank and falls back to the
  real ``energy_fn.with_params(opt_params)`` for anything else a loss function may ask of it.

If anything in the chain cannot be lowered (an op outside the small set below, data-dependent control flow, a
composition that does not fuse into one launch group) ``bind`` simply returns ``energy_fn.with_params(opt_params)``.
"""

from __future__ import annotations

import ctypes as C
import dataclasses as dc
import operator
import threading
import warnings
import weakref
from collections.abc import Mapping

import numpy as np
import torch

from mythos_b200 import _lib

OP_CONST, OP_INPUT, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_NEG, OP_RECIP, OP_EXP, OP_LOG, OP_SQRT, OP_POW = range(12)


class _CTape(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("n_inputs", C.c_int32), ("n_outputs", C.c_int32), ("_pad", C.c_int32),
                ("op", C.c_void_p), ("arg0", C.c_void_p), ("arg1", C.c_void_p), ("imm", C.c_void_p), ("out", C.c_void_p)]


class LoweringError(RuntimeError):
    pass


class _Builder:
    """Scalar SSA program under construction (hash-consed, constants folded)."""

    def __init__(self):
        self.op: list[int] = []
        self.a0: list[int] = []
        self.a1: list[int] = []
        self.imm: list[float] = []
        self._memo: dict = {}

    def _emit(self, op, a=-1, b=-1, imm=0.0) -> int:
        key = (op, a, b, float(imm).hex())
        hit = self._memo.get(key)
        if hit is not None:
            return hit
        self.op.append(op), self.a0.append(a), self.a1.append(b), self.imm.append(float(imm))
        self._memo[key] = len(self.op) - 1
        return len(self.op) - 1

    def const(self, v) -> int:
        return self._emit(OP_CONST, imm=float(v))

    def input(self, k) -> int:
        return self._emit(OP_INPUT, a=int(k))

    def _cval(self, k):
        return self.imm[k] if self.op[k] == OP_CONST else None

    def binary(self, op, a, b) -> int:
        ca, cb = self._cval(a), self._cval(b)
        if ca is not None and cb is not None:
            with np.errstate(all="ignore"):
                x, y = np.float64(ca), np.float64(cb)
                return self.const({OP_ADD: x + y, OP_SUB: x - y, OP_MUL: x * y, OP_DIV: x / y}[op])
        return self._emit(op, a, b)

    def unary(self, op, a, imm=0.0) -> int:
        ca = self._cval(a)
        if ca is not None:
            with np.errstate(all="ignore"):
                x = np.float64(ca)
                val = {OP_NEG: lambda: -x, OP_RECIP: lambda: 1.0 / x, OP_EXP: lambda: np.exp(x), OP_LOG: lambda: np.log(x),
                       OP_SQRT: lambda: np.sqrt(x), OP_POW: lambda: x * x if imm == 2.0 else np.power(x, imm)}[op]()
                return self.const(val)
        return self._emit(op, a, -1, imm)


def _lower(graph: torch.fx.GraphModule, n_inputs: int):
    """FX graph of small float64 tensor ops -> (op, arg0, arg1, imm, out) numpy arrays."""
    B = _Builder()
    env: dict = {}
    aten = torch.ops.aten

    def arr(x):
        """Operand -> int64 ndarray of node ids."""
        if isinstance(x, np.ndarray):
            return x
        if isinstance(x, torch.fx.Node):
            v = env[x]
            if not isinstance(v, np.ndarray):
                raise LoweringError(f"non-tensor operand {x}")
            return v
        if isinstance(x, (int, float, bool)):
            return np.array(B.const(float(x)), dtype=np.int64)
        if isinstance(x, torch.Tensor):
            return const_tensor(x)
        raise LoweringError(f"operand {x!r}")

    def const_tensor(t: torch.Tensor):
        vals = t.detach().to(torch.float64).cpu().numpy()
        out = np.empty(vals.shape, dtype=np.int64)
        for idx in np.ndindex(vals.shape):
            out[idx] = B.const(vals[idx])
        return out if vals.shape else np.array(B.const(float(vals)), dtype=np.int64)

    def ew2(op, a, b):
        a, b = np.broadcast_arrays(arr(a), arr(b))
        out = np.empty(a.shape, dtype=np.int64)
        for idx in np.ndindex(a.shape):
            out[idx] = B.binary(op, int(a[idx]), int(b[idx]))
        return out

    def ew1(op, a, imm=0.0):
        a = arr(a)
        out = np.empty(a.shape, dtype=np.int64)
        for idx in np.ndindex(a.shape):
            out[idx] = B.unary(op, int(a[idx]), imm)
        return out

    def scaled(x, alpha):
        return x if alpha == 1 else ew2(OP_MUL, x, alpha)

    outputs = None
    for node in graph.graph.nodes:
        if node.op == "placeholder":
            env[node] = np.array([B.input(k) for k in range(n_inputs)], dtype=np.int64)
        elif node.op == "get_attr":
            env[node] = const_tensor(getattr(graph, node.target))
        elif node.op == "output":
            outputs = node.args[0]
        elif node.op == "call_function":
            t, a, kw = node.target, node.args, node.kwargs
            if t is operator.getitem:
                env[node] = env[a[0]][a[1]]
            elif t == aten.select.int:
                env[node] = np.asarray(np.take(arr(a[0]), a[2], axis=a[1]))
            elif t == aten.slice.Tensor:
                x = arr(a[0])
                dim = a[1] if len(a) > 1 else 0
                start = a[2] if len(a) > 2 and a[2] is not None else 0
                end = a[3] if len(a) > 3 and a[3] is not None else x.shape[dim]
                step = a[4] if len(a) > 4 else 1
                sl = [slice(None)] * x.ndim
                sl[dim] = slice(start, min(end, x.shape[dim]), step)
                env[node] = x[tuple(sl)]
            elif t in (aten.view.default, aten.reshape.default, aten._unsafe_view.default):
                env[node] = arr(a[0]).reshape(tuple(a[1]))
            elif t == aten.expand.default:
                x = arr(a[0])
                shape = tuple(x.shape[k - (len(a[1]) - x.ndim)] if s == -1 else s for k, s in enumerate(a[1]))
                env[node] = np.broadcast_to(x, shape)
            elif t == aten.unsqueeze.default:
                env[node] = np.expand_dims(arr(a[0]), a[1])
            elif t in (aten.squeeze.dim, aten.squeeze.dims):
                env[node] = np.squeeze(arr(a[0]), axis=tuple(a[1]) if isinstance(a[1], (list, tuple)) else a[1])
            elif t == aten.squeeze.default:
                env[node] = np.squeeze(arr(a[0]))
            elif t == aten.t.default:
                env[node] = arr(a[0]).T
            elif t == aten.transpose.int:
                env[node] = np.swapaxes(arr(a[0]), a[1], a[2])
            elif t == aten.permute.default:
                env[node] = np.transpose(arr(a[0]), tuple(a[1]))
            elif t == aten.stack.default:
                env[node] = np.stack([arr(x) for x in a[0]], axis=a[1] if len(a) > 1 else 0)
            elif t == aten.cat.default:
                env[node] = np.concatenate([arr(x) for x in a[0]], axis=a[1] if len(a) > 1 else 0)
            elif t == aten.unbind.int:
                x = arr(a[0])
                dim = a[1] if len(a) > 1 else 0
                env[node] = tuple(np.asarray(np.take(x, k, axis=dim)) for k in range(x.shape[dim]))
            elif t in (aten.add.Tensor, aten.add.Scalar):
                env[node] = ew2(OP_ADD, a[0], scaled(a[1], kw.get("alpha", 1)))
            elif t in (aten.sub.Tensor, aten.sub.Scalar):
                env[node] = ew2(OP_SUB, a[0], scaled(a[1], kw.get("alpha", 1)))
            elif t in (aten.rsub.Scalar, aten.rsub.Tensor):
                env[node] = ew2(OP_SUB, a[1], scaled(a[0], kw.get("alpha", 1)))
            elif t in (aten.mul.Tensor, aten.mul.Scalar):
                env[node] = ew2(OP_MUL, a[0], a[1])
            elif t in (aten.div.Tensor, aten.div.Scalar):
                if kw.get("rounding_mode") is not None:
                    raise LoweringError("div with rounding_mode")
                env[node] = ew2(OP_DIV, a[0], a[1])
            elif t == aten.pow.Tensor_Scalar:
                env[node] = ew1(OP_POW, a[0], float(a[1]))
            elif t == aten.neg.default:
                env[node] = ew1(OP_NEG, a[0])
            elif t == aten.reciprocal.default:
                env[node] = ew1(OP_RECIP, a[0])
            elif t == aten.exp.default:
                env[node] = ew1(OP_EXP, a[0])
            elif t == aten.log.default:
                env[node] = ew1(OP_LOG, a[0])
            elif t == aten.sqrt.default:
                env[node] = ew1(OP_SQRT, a[0])
            elif t == aten.rsqrt.default:
                env[node] = ew1(OP_RECIP, ew1(OP_SQRT, a[0]))
            elif t in (aten.zeros.default, aten.ones.default):
                fill = 0.0 if t == aten.zeros.default else 1.0
                env[node] = np.full(tuple(a[0]), B.const(fill), dtype=np.int64)
            elif t == aten.full.default:
                env[node] = np.full(tuple(a[0]), B.const(float(a[1])), dtype=np.int64)
            elif t == aten.scalar_tensor.default:
                env[node] = np.array(B.const(float(a[0])), dtype=np.int64)
            elif t in (aten.clone.default, aten.detach.default, aten.alias.default, aten.lift_fresh_copy.default,
                       aten._to_copy.default, aten.contiguous.default):
                dt = kw.get("dtype")
                if dt is not None and dt != torch.float64:
                    raise LoweringError(f"cast to {dt} inside the theta chain")
                env[node] = arr(a[0])
            elif t == aten.sum.default:
                x = arr(a[0]).reshape(-1)
                acc = int(x[0])
                for k in x[1:]:
                    acc = B.binary(OP_ADD, acc, int(k))
                env[node] = np.array(acc, dtype=np.int64)
            else:
                raise LoweringError(f"unsupported op in the theta chain: {t}")
        else:
            raise LoweringError(f"unsupported node kind {node.op}")
    if isinstance(outputs, (tuple, list)):
        if len(outputs) != 1:
            raise LoweringError("expected one output")
        outputs = outputs[0]
    out = arr(outputs).reshape(-1)

    # dead-node elimination + renumbering (operands always precede their use)
    n = len(B.op)
    live = np.zeros(n, dtype=bool)
    live[out] = True
    for k in range(n - 1, -1, -1):
        if live[k]:
            if B.a0[k] >= 0 and B.op[k] != OP_INPUT:
                live[B.a0[k]] = True
            if B.a1[k] >= 0:
                live[B.a1[k]] = True
    new = np.cumsum(live) - 1
    keep = np.nonzero(live)[0]
    op = np.array([B.op[k] for k in keep], dtype=np.int32)
    a0 = np.array([(B.a0[k] if B.op[k] == OP_INPUT else (new[B.a0[k]] if B.a0[k] >= 0 else -1)) for k in keep], dtype=np.int32)
    a1 = np.array([(new[B.a1[k]] if B.a1[k] >= 0 else -1) for k in keep], dtype=np.int32)
    imm = np.array([B.imm[k] for k in keep], dtype=np.float64)
    return op, a0, a1, imm, new[out].astype(np.int32)


class ThetaTape:
    """A lowered chain: ``forward(theta (n_in,)) -> bank (n_out,)`` and its VJP, replayed by the C library."""

    def __init__(self, op, a0, a1, imm, out, n_inputs: int):
        self.op, self.a0, self.a1, self.imm, self.out = (np.ascontiguousarray(x) for x in (op, a0, a1, imm, out))
        self.n_inputs, self.n_nodes, self.n_outputs = int(n_inputs), int(len(op)), int(len(out))
        self._c = _CTape(self.n_nodes, self.n_inputs, self.n_outputs, 0, self.op.ctypes.data, self.a0.ctypes.data,
                         self.a1.ctypes.data, self.imm.ctypes.data, self.out.ctypes.data)
        lib = _lib.lib()
        self._fwd, self._vjp = lib.mythos_b200_theta_tape_forward, lib.mythos_b200_theta_tape_vjp

    def forward(self, x: np.ndarray):
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.shape != (self.n_inputs,):
            raise _lib.MythosB200Error(f"theta tape expects {self.n_inputs} inputs, got {x.shape}")
        vals = np.empty(self.n_nodes, dtype=np.float64)
        out = np.empty(self.n_outputs, dtype=np.float64)
        _lib.check(self._fwd(C.byref(self._c), x.ctypes.data, vals.ctypes.data, out.ctypes.data), "theta_tape_forward")
        return out, vals

    def vjp(self, vals: np.ndarray, g: np.ndarray) -> np.ndarray:
        g = np.ascontiguousarray(g, dtype=np.float64)
        adj = np.empty(self.n_nodes, dtype=np.float64)
        grad = np.empty(self.n_inputs, dtype=np.float64)
        _lib.check(self._vjp(C.byref(self._c), vals.ctypes.data, g.ctypes.data, adj.ctypes.data, grad.ctypes.data), "theta_tape_vjp")
        return grad


class _Replay(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vec, tape):
        out, vals = tape.forward(vec.detach().cpu().numpy())
        ctx.tape, ctx.vals, ctx.dev = tape, vals, vec.device
        return torch.from_numpy(out)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        grad = ctx.tape.vjp(ctx.vals, g.detach().cpu().numpy())
        return torch.from_numpy(grad).to(ctx.dev), None


# ------------------------------------------------------------------------------------------------- per-function cache
@dc.dataclass
class _Entry:
    ref: weakref.ref
    names: tuple
    shapes: tuple
    tape: ThetaTape | None  # None: this (function, parameter set) cannot take the fast path
    plan: object = None
    weights: torch.Tensor | None = None


_CACHE: dict = {}
_LOCK = threading.Lock()


class FlatParams(Mapping):
    """A parameter dict backed by ONE flat float64 vector: ``params[name]`` is a view of it.  ``compute_loss_and_grad``
    differentiates with respect to the vector (one autograd leaf instead of a hundred) and ``bind`` feeds it to the tape
    without re-assembling it.  Behaves like the dict the reference hands to ``loss_fn`` (keys, items, ``**`` unpacking)."""

    def __init__(self, source: dict, flat: torch.Tensor | None = None):
        # (built once per optimiser step on the critical path in front of the first kernel launch: python floats / one
        # torch.tensor call instead of a hundred as_tensor + reshape + cat calls, 0.28 -> 0.05 ms for 103 parameters)
        self.names = tuple(sorted(source))
        if flat is None and source:
            # the common case (every parameter a float64 scalar tensor on the host): one stack, checked on its result
            try:
                stacked = torch.stack([source[k] for k in self.names])
            except (TypeError, RuntimeError):  # python floats, tables, mixed shapes or devices: the general path below
                stacked = None
            if stacked is not None and stacked.dim() == 1 and stacked.dtype == torch.float64 and not stacked.is_cuda:
                n = len(self.names)
                self.shapes, self.sizes = ((),) * n, (1,) * n
                self.offsets = dict(zip(self.names, range(n)))
                self.flat = stacked.detach()
                self._views = {}
                return
        shapes, sizes, data = [], [], []
        for k in self.names:
            v = source[k]
            if isinstance(v, torch.Tensor):
                shapes.append(tuple(v.shape))
                sizes.append(v.numel())
                if flat is None:
                    if v.numel() == 1:
                        data.append(v.item())
                    else:
                        data.extend(v.detach().reshape(-1).tolist())
            else:
                shapes.append(())
                sizes.append(1)
                if flat is None:
                    data.append(float(v))
        self.shapes, self.sizes = tuple(shapes), tuple(sizes)
        self.offsets = {}
        off = 0
        for k, sz in zip(self.names, self.sizes):
            self.offsets[k] = off
            off += sz
        if flat is None:
            flat = torch.tensor(data, dtype=torch.float64)
        self.flat = flat
        self._views: dict = {}

    def like(self, flat: torch.Tensor) -> "FlatParams":
        new = object.__new__(FlatParams)
        new.names, new.shapes, new.sizes, new.offsets, new.flat, new._views = self.names, self.shapes, self.sizes, self.offsets, flat, {}
        return new

    def __getitem__(self, key):
        v = self._views.get(key)
        if v is None:
            k = self.names.index(key) if key in self.offsets else None
            if k is None:
                raise KeyError(key)
            off, shp = self.offsets[key], self.shapes[k]
            v = self._views[key] = self.flat[off] if not shp else self.flat[off:off + self.sizes[k]].reshape(shp)
        return v

    def __iter__(self):
        return iter(self.names)

    def __len__(self):
        return len(self.names)

    def unflatten(self, vec: torch.Tensor) -> dict:
        if all(not shp for shp in self.shapes):  # all scalars: one unbind instead of a view per parameter
            return dict(zip(self.names, vec.unbind(0)))
        parts = torch.split(vec, list(self.sizes))
        return {k: (p.reshape(()) if not shp else p.reshape(shp)) for k, shp, p in zip(self.names, self.shapes, parts)}


def _flatten(opt_params, names, shapes) -> torch.Tensor:
    if isinstance(opt_params, FlatParams) and opt_params.names == names and opt_params.shapes == shapes:
        return opt_params.flat
    parts = []
    for k, shp in zip(names, shapes):
        v = torch.as_tensor(opt_params[k], dtype=torch.float64)
        if tuple(v.shape) != shp:
            raise _lib.MythosB200Error(f"parameter {k!r} changed shape: {tuple(v.shape)} vs {shp}")
        parts.append(v.reshape(-1))
    return torch.cat(parts) if parts else torch.zeros(0, dtype=torch.float64)


def _chain_fn(energy_fn, names, shapes):
    from mythos_b200.energy import model as kmodel

    sizes = [int(np.prod(s)) if s else 1 for s in shapes]

    def chain(vec: torch.Tensor) -> torch.Tensor:
        d, off = {}, 0
        for k, shp, sz in zip(names, shapes, sizes):
            d[k] = vec[off] if not shp else vec[off:off + sz].reshape(shp)
            off += sz
        fns = energy_fn.with_params(d).energy_fns
        return kmodel.bank_vector(fns, any(fn.HYBRID for fn in fns))

    return chain


def _build_entry(energy_fn, opt_params, names, shapes) -> _Entry:
    from torch.fx.experimental.proxy_tensor import make_fx

    from mythos_b200.energy import model as kmodel

    entry = _Entry(ref=weakref.ref(energy_fn), names=names, shapes=shapes, tape=None)
    fns = getattr(energy_fn, "energy_fns", None)
    if not fns or len(kmodel.fusable_groups(fns)) != 1:
        return entry
    if any("pseq" in fn.params and fn.params.pseq is not None for fn in fns) or any(k in ("pseq", "pseq_constraints") for k in names):
        return entry  # probabilistic sequences chain through the configurations themselves (mythos_b200.energy.pseq)
    try:
        chain = _chain_fn(energy_fn, names, shapes)
        vec = _flatten(opt_params, names, shapes).detach().cpu()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            graph = make_fx(chain)(vec.clone())
        tape = ThetaTape(*_lower(graph, vec.numel()), n_inputs=vec.numel())
        # the lowering is only trusted after it reproduces the eager chain: values and a vector-Jacobian product
        probe = vec.clone().requires_grad_(True)
        want = chain(probe)
        cot = torch.linspace(0.5, 1.5, want.numel(), dtype=torch.float64)
        (want_g,) = torch.autograd.grad((want * cot).sum(), probe, allow_unused=True)
        want_g = torch.zeros_like(vec) if want_g is None else want_g
        got, vals = tape.forward(vec.numpy())
        got_g = tape.vjp(vals, cot.numpy())
        scale = lambda t: 1e-12 * (1.0 + float(t.abs().max()))  # noqa: E731
        if not (np.allclose(got, want.detach().numpy(), rtol=1e-12, atol=scale(want.detach()))
                and np.allclose(got_g, want_g.numpy(), rtol=1e-10, atol=1e-10 * (1.0 + float(want_g.abs().max())))):
            raise LoweringError("replayed chain does not reproduce the eager chain")
    except (LoweringError, RuntimeError, TypeError, ValueError, KeyError, IndexError, NotImplementedError) as err:
        if isinstance(err, _lib.MythosB200Error):
            raise
        warnings.warn(f"theta chain of {type(energy_fn).__name__} is evaluated eagerly ({err})", stacklevel=3)
        return entry
    entry.tape = tape
    entry.plan = kmodel.plan_for(fns)
    w = torch.zeros(_lib.N_TERMS, dtype=torch.float64)
    wt = torch.ones(len(fns), dtype=torch.float64) if energy_fn.weights is None else torch.as_tensor(energy_fn.weights, dtype=torch.float64).cpu()
    for k, fn in enumerate(fns):
        w[fn.TERM] = wt[k]
    entry.weights = w
    return entry


def _entry_for(energy_fn, opt_params) -> _Entry | None:
    if isinstance(opt_params, FlatParams):
        names, shapes = opt_params.names, opt_params.shapes
    else:
        names = tuple(sorted(opt_params))
        try:
            shapes = tuple(tuple(torch.as_tensor(opt_params[k]).shape) for k in names)
        except (TypeError, ValueError, RuntimeError):
            return None
    key = (id(energy_fn), names, shapes)
    with _LOCK:
        hit = _CACHE.get(key)
        if hit is not None and hit.ref() is energy_fn:
            return hit
    with torch.enable_grad():  # (the first call may come from inside a no_grad block; the self-check differentiates)
        entry = _build_entry(energy_fn, opt_params, names, shapes)
    with _LOCK:
        for k in [k for k, e in _CACHE.items() if e.ref() is None]:
            del _CACHE[k]
        _CACHE[key] = entry
    return entry


class BoundEnergyFunction:
    """``energy_fn.with_params(opt_params)`` whose ``map`` runs on the replayed parameter bank.  Everything else
    (``__call__``, ``compute_terms``, ``params_dict`` ...) is answered by the real ``with_params`` result, built on first use."""

    def __init__(self, energy_fn, opt_params, entry: _Entry, bank: torch.Tensor):
        object.__setattr__(self, "_base", energy_fn)
        object.__setattr__(self, "_opt", opt_params)
        object.__setattr__(self, "_entry", entry)
        object.__setattr__(self, "_bank", bank)
        object.__setattr__(self, "_real", None)

    def _materialise(self):
        if self._real is None:
            object.__setattr__(self, "_real", self._base.with_params(self._opt))
        return self._real

    def map(self, body_sequence, observables=None):
        from mythos_b200.energy import model as kmodel
        from mythos_b200.energy.base import _frames

        c, q, _ = _frames(body_sequence)
        plan = dc.replace(self._entry.plan, bank=self._bank)
        return kmodel.evaluate_with_observables(plan, c, q, self._entry.weights, observables)

    def with_params(self, *repl_dicts, **repl_kwargs):
        return self._materialise().with_params(*repl_dicts, **repl_kwargs)

    def __call__(self, *args, **kwargs):
        return self._materialise()(*args, **kwargs)

    def __getattr__(self, name):
        return getattr(self._materialise(), name)

    def __setattr__(self, name, value):
        raise AttributeError("BoundEnergyFunction is frozen")


ENABLED = True  # tests switch the replay off to compare it with the eager chain


def bind(energy_fn, opt_params):
    """``energy_fn.with_params(opt_params)``, through the replayed chain when this function / parameter set allows it."""
    if not ENABLED:
        return energy_fn.with_params(opt_params)
    if not isinstance(opt_params, Mapping) or not opt_params or not hasattr(energy_fn, "energy_fns"):
        return energy_fn.with_params(opt_params)
    entry = _entry_for(energy_fn, opt_params)
    if entry is None or entry.tape is None:
        return energy_fn.with_params(opt_params)
    vec = _flatten(opt_params, entry.names, entry.shapes)
    return BoundEnergyFunction(energy_fn, opt_params, entry, _Replay.apply(vec, entry.tape))


def clear_cache() -> None:
    with _LOCK:
        _CACHE.clear()
