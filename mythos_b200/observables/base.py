"""Observables: base class, quartet helper and the kernel glue (interface of ``mythos/observables/base.py:14-66``).

The reference computes an observable by re-deriving every interaction site of every frame with
``jax.vmap(rigid_body_transform_fn)(trajectory)`` and then mapping a per-base-pair function over it.  Here the four
structural observables of a duplex -- propeller twist, rise, pitch angle, diameter -- are columns of ONE kernel result
``(F, 4)`` (``csrc/observables_dev.cuh``): either the epilogue of the frame-resident energy kernel of a DiffTRe pass (the
frame is already in shared memory; ``ObservableSet`` + ``energy_fn.map(states, observables=...)``), or a standalone
launch that reads only the listed nucleotides.  Results are remembered per trajectory tensor, so the observable a loss
function asks for right after the energy pass costs no second pass over the frames.
"""

from __future__ import annotations

import ctypes as C
import dataclasses as dc
import itertools
import threading
import weakref
from collections.abc import Callable

import numpy as np
import torch

from mythos_b200 import _lib, space

ERR_RIGID_BODY_TRANSFORM_FN_REQUIRED = "rigid_body_transform_fn must be provided"


@dc.dataclass(frozen=True)
class BaseObservable:
    """Base class for observables."""

    rigid_body_transform_fn: Callable

    def __call__(self, trajectory) -> torch.Tensor:
        """Calculate the observable."""


def get_duplex_quartets(n_nucs_per_strand: int) -> torch.Tensor:
    """All quartets (pairs of adjacent base pairs) of a duplex, ``(n-1, 2, 2)`` int32 (``base.py:47-66``)."""
    s1 = list(range(n_nucs_per_strand))
    s2 = list(range(n_nucs_per_strand, n_nucs_per_strand * 2))
    s2.reverse()
    bps = list(zip(s1, s2, strict=True))
    return torch.tensor(list(map(list, itertools.pairwise(bps))), dtype=torch.int32).reshape(-1, 2, 2)


# ------------------------------------------------------------------------------------------------- kernel glue
def _index_key(idx) -> tuple | None:
    if idx is None:
        return None
    a = np.ascontiguousarray(np.asarray(idx.cpu() if isinstance(idx, torch.Tensor) else idx, dtype=np.int32))
    return (a.shape, a.tobytes())


@dc.dataclass
class ObservableRequest:
    """What one kernel evaluation needs: the index lists on the device and where the ``(F, 4)`` result goes."""

    base_pairs: torch.Tensor | None  # (P,2) int32 device
    quartets: torch.Tensor | None  # (Q,4) int32 device
    sigma_backbone: float
    key: tuple  # identity of (geometry, box, lists, sigma) for the result cache
    out: torch.Tensor | None = None

    def struct(self) -> _lib.ObservableSpec:
        s = _lib.ObservableSpec()
        s.base_pairs = _lib.ptr(self.base_pairs)
        s.n_base_pairs = 0 if self.base_pairs is None else int(self.base_pairs.shape[0])
        s.quartets = _lib.ptr(self.quartets)
        s.n_quartets = 0 if self.quartets is None else int(self.quartets.shape[0])
        s.sigma_backbone = float(self.sigma_backbone)
        return s


def make_request(model_key, base_pairs, quartets, sigma_backbone, device) -> ObservableRequest:
    bp = None
    if base_pairs is not None:
        bp = torch.as_tensor(base_pairs).to(device=device, dtype=torch.int32).reshape(-1, 2).contiguous()
    qt = None
    if quartets is not None:
        qt = torch.as_tensor(quartets).to(device=device, dtype=torch.int32).reshape(-1, 4).contiguous()
    return ObservableRequest(bp, qt, float(sigma_backbone), (model_key, _index_key(base_pairs), _index_key(quartets), float(sigma_backbone)))


class _ColumnCache:
    """(F,4) observable columns per (frames tensor, request key); entries die with the frames tensor or when it is
    written in place.  Filled by the standalone launch and by the fused epilogue of the energy pass."""

    def __init__(self):
        self._lock = threading.Lock()
        self._data: dict = {}

    def get(self, center: torch.Tensor, key):
        with self._lock:
            e = self._data.get(id(center))
            if e is None or e["ref"]() is not center or e["version"] != center._version:
                self._data.pop(id(center), None)
                return None
            return e["cols"].get(key)

    def put(self, center: torch.Tensor, key, cols: torch.Tensor) -> None:
        with self._lock:
            e = self._data.get(id(center))
            if e is None or e["ref"]() is not center or e["version"] != center._version:
                k = id(center)
                e = {"ref": weakref.ref(center, lambda _r, k=k: self._drop(k)), "version": center._version, "cols": {}}
                self._data[k] = e
            if len(e["cols"]) > 16:
                e["cols"].pop(next(iter(e["cols"])))
            e["cols"][key] = cols

    def _drop(self, k) -> None:
        with self._lock:
            self._data.pop(k, None)

    def clear(self) -> None:
        with self._lock:
            self._data.clear()


COLUMNS = _ColumnCache()


def model_of(transform_fn, displacement_fn) -> tuple[_lib.Model, tuple]:
    """Kernel model description (flavour geometry + box) of an observable's transform / displacement functions."""
    from mythos_b200.energy import model as kmodel

    kind, geoms = kmodel.geometry_of(transform_fn)
    if kind == "na1":
        raise NotImplementedError("observables of hybrid (NA1) trajectories: pass the DNA or RNA nucleotide transform_fn")
    m = _lib.Model()
    m.n_banks = 1
    m.geom[0] = geoms[0]
    box = space.box_of(displacement_fn) if displacement_fn is not None else (0.0, 0.0, 0.0)
    for d in range(3):
        m.box[d] = box[d]
    g = geoms[0]
    return m, (kind, tuple(g.back), float(g.base), tuple(box))


def launch(model: _lib.Model, center: torch.Tensor, quat: torch.Tensor, req: ObservableRequest) -> torch.Tensor:
    """Standalone kernel: ``(F,N,3), (F,N,4)`` on the device -> ``(F, 4)`` columns (propeller, rise, pitch angle, diameter)."""
    _lib.require_cuda(center, "trajectory.center")
    F, n = center.shape[0], center.shape[1]
    out = torch.empty((F, _lib.N_OBS), dtype=center.dtype, device=center.device)
    fn = getattr(_lib.lib(), f"mythos_b200_observables_{_lib.suffix(center.dtype)}")
    spec = req.struct()
    with torch.cuda.device(center.device):
        _lib.check(fn(_lib.current_stream(center.device), C.pointer(model), n, F, center.contiguous().data_ptr(),
                      quat.to(center.dtype).contiguous().data_ptr(), None, C.byref(spec), out.data_ptr()), "mythos_b200_observables")
    return out


def columns(transform_fn, displacement_fn, trajectory, base_pairs=None, quartets=None, sigma_backbone: float = 0.0) -> torch.Tensor:
    """``(F, 4)`` observable columns of a trajectory, from the cache (fused epilogue of the last energy pass over these
    frames, or an earlier call) or from one standalone launch."""
    center = trajectory.center
    quat = trajectory.orientation.vec if hasattr(trajectory.orientation, "vec") else trajectory.orientation
    single = center.dim() == 2
    model, mkey = model_of(transform_fn, displacement_fn)
    req = make_request(mkey, base_pairs, quartets, sigma_backbone, center.device)
    if not single:
        hit = COLUMNS.get(center, req.key)
        if hit is not None:
            return hit
    cols = launch(model, center.unsqueeze(0) if single else center, quat.unsqueeze(0) if single else quat, req)
    if single:
        return cols
    COLUMNS.put(center, req.key, cols)
    return cols


def _covers(have, want) -> bool:
    return want is None or have == want


@dc.dataclass(frozen=True)
class ObservableSet:
    """Observables to evaluate INSIDE the energy pass over the same frames (``energy_fn.map(states, observables=...)`` or
    ``loss_fn.fused_observables = ObservableSet([...])`` for ``compute_loss``): they must agree on the base-pair list, the
    quartet list, the nucleotide geometry and the box, because one epilogue serves all of them."""

    observables: tuple

    def __init__(self, observables):
        object.__setattr__(self, "observables", tuple(observables))
        if not self.observables:
            raise ValueError("ObservableSet needs at least one observable")

    def request(self, device) -> ObservableRequest:
        tf = self.observables[0].rigid_body_transform_fn
        disp = next((getattr(o, "displacement_fn", None) for o in self.observables if getattr(o, "displacement_fn", None) is not None), None)
        _, mkey = model_of(tf, disp)
        bp = qt = None
        sigma = 0.0
        for o in self.observables:
            _, k = model_of(o.rigid_body_transform_fn, getattr(o, "displacement_fn", None) or disp)
            if k != mkey:
                raise ValueError("fused observables must share nucleotide geometry and box")
            obp, oqt = getattr(o, "h_bonded_base_pairs", None), getattr(o, "quartets", None)
            if obp is not None:
                if bp is not None and _index_key(bp) != _index_key(obp):
                    raise ValueError("fused observables must share one base-pair list")
                bp = obp
            if oqt is not None:
                if qt is not None and _index_key(qt) != _index_key(oqt):
                    raise ValueError("fused observables must share one quartet list")
                qt = oqt
            sigma = float(getattr(o, "sigma_backbone", sigma) or sigma)
        return make_request(mkey, bp, qt, sigma, device)

    def publish(self, center: torch.Tensor, req: ObservableRequest) -> None:
        """Make the fused result visible to the member observables' ``__call__`` on the same trajectory."""
        if req.out is None:
            return
        for o in self.observables:
            obp, oqt = getattr(o, "h_bonded_base_pairs", None), getattr(o, "quartets", None)
            sig = float(getattr(o, "sigma_backbone", 0.0) or 0.0)
            own = make_request(req.key[0], obp, oqt, sig, center.device)
            COLUMNS.put(center, own.key, req.out)
