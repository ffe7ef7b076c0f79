"""From a list of energy-term objects to one fused kernel launch: model description, parameter bank, device inputs.

The reference gives every term its own traced sub-graph (``mythos/energy/base.py:312-314``).  Here the terms of a
composition are folded into a ``Plan``:

* ``mb_model`` -- flavour geometry (from the term's ``transform_fn``), per-bank functional forms (from the term
  classes), box (from ``displacement_fn``), half-charged-ends flag (from the Debye configuration);
* the kernel-level parameter bank(s) -- every configuration packed by name into the ``MB_PARAM_LIST`` order; the
  pack is a ``torch.stack`` of the configuration's scalars so autograd chains d/dparams back to theta;
* int32 device copies of ``seq``, bonds, ``nt_type``, ``is_end`` and the unbonded pair list (cached per device).
"""

from __future__ import annotations

import dataclasses as dc
import re
import threading
from functools import lru_cache
from typing import Any

import torch

from mythos_b200 import _lib, space
from mythos_b200.energy import functional
from mythos_b200.energy.base_smoothing_functions import as_t

STACK_DNA, STACK_RNA = 0, 1
CROSS_DNA1, CROSS_RNA2 = 0, 1
COAX_DNA1, COAX_DNA2 = 0, 1
BANKS = ("dna", "rna", "drh")


# --------------------------------------------------------------------------------------------- bank layout
@lru_cache(maxsize=1)
def bank_layout() -> dict[str, list[tuple[int, str, tuple[int, int] | None]]]:
    """term -> [(bank index, configuration field, (row, col) or None)] parsed from the library's name table."""
    out: dict[str, list] = {}
    pat = re.compile(r"^([a-z_]+)\.([A-Za-z0-9_]+)(?:\[(\d),(\d)\])?$")
    for idx, name in enumerate(_lib.param_names()):
        m = pat.match(name)
        if not m:
            raise _lib.MythosB200Error(f"unparseable kernel parameter name {name!r}")
        sub = (int(m.group(3)), int(m.group(4))) if m.group(3) is not None else None
        out.setdefault(m.group(1), []).append((idx, m.group(2), sub))
    return out


def pack_bank(configs: list[Any]) -> torch.Tensor:
    """One parameter bank (P,) float64 on the host from initialised term configurations (missing slots = 0)."""
    P = _lib.param_count()
    zero = torch.zeros((), dtype=torch.float64)
    vals = [zero] * P
    layout = bank_layout()
    for cfg in configs:
        for idx, field, sub in layout.get(cfg.term, ()):
            v = getattr(cfg, field, None) if field in cfg else None
            if v is None:
                continue
            v = as_t(v)
            if v.dtype != torch.float64:
                v = v.to(torch.float64)
            # (no-op views would each leave a node in the autograd graph of the theta -> bank chain)
            vals[idx] = v[sub] if sub is not None else (v if v.dim() == 0 else v.reshape(()))
    return torch.stack(vals)


def bank_vector(fns: list, hybrid: bool) -> torch.Tensor:
    """(n_banks*P,) float64 host vector of a fusable group of terms (NA1: the DNA, RNA and hybrid banks in that order)."""
    if not hybrid:
        return pack_bank([fn.params for fn in fns])
    banks = []
    for b in BANKS:
        cfgs = [getattr(fn.params, f"{b}_config") for fn in fns if f"{b}_config" in fn.params]
        banks.append(pack_bank([c for c in cfgs if c is not None]))
    return torch.cat(banks)


# --------------------------------------------------------------------------------------------- geometry
def geometry_of(transform_fn) -> tuple[str, list[_lib.FlavourGeom]]:
    """(kind, flavour geometries) of a ``functools.partial(Nucleotide.from_rigid_body, **constants)``."""
    func = getattr(transform_fn, "func", transform_fn)
    cls = getattr(func, "nucleotide_cls", None)
    if cls is None:
        raise _lib.MythosB200Error(
            "transform_fn must be functools.partial(<Nucleotide>.from_rigid_body, ...) of a mythos_b200 nucleotide class"
        )
    kw = dict(getattr(transform_fn, "keywords", {}) or {})
    return cls.KIND, cls.kernel_geometry(**{k: float(as_t(v)) for k, v in kw.items()})


# --------------------------------------------------------------------------------------------- caches
_TOPO_CACHE: dict[tuple, tuple[tuple, Any]] = {}  # keys carry the device: per-device entries
_TOPO_LOCK = threading.Lock()  # evaluations may come from several host threads (XLA executor threads, user threads)


def _cached(key: tuple, keep: tuple, make):
    with _TOPO_LOCK:
        hit = _TOPO_CACHE.get(key)
        if hit is not None and all(a is b for a, b in zip(hit[0], keep)):
            return hit[1]
    val = make()  # (device copies are made outside the lock; a racing duplicate is harmless, the last one is kept)
    with _TOPO_LOCK:
        if len(_TOPO_CACHE) > 64:
            _TOPO_CACHE.pop(next(iter(_TOPO_CACHE)))
        _TOPO_CACHE[key] = (keep, val)
    return val


def check_pair_layout(shape) -> None:
    """The kernels take the OrderedSparse layout (2,U) / (F,2,U) -- what the reference hands its energy functions after
    transposing ``topology.unbonded_neighbors (U,2)`` (``energy/base.py:138``, ``jaxmd.py:78,86``).  A (U,2) array would be
    read as a two-entry list, silently: refuse it."""
    shape = tuple(shape)
    if len(shape) not in (2, 3) or shape[-2] != 2:
        hint = " -- this looks like (U,2) `topology.unbonded_neighbors`: pass its transpose" if len(shape) == 2 and shape[-1] == 2 else ""
        raise _lib.MythosB200Error(f"unbonded_neighbors must have shape (2,U) or (F,2,U), got {shape}{hint}")


def device_pairs(pairs, device) -> torch.Tensor | None:
    """(2,U) or (F,2,U) int32 on the device; accepts numpy / lists / tensors; a (U,2) array is refused, not transposed."""
    if pairs is None:
        return None
    check_pair_layout(getattr(pairs, "shape", None) if hasattr(pairs, "shape") else torch.as_tensor(pairs).shape)
    if isinstance(pairs, torch.Tensor) and pairs.device == device and pairs.dtype == torch.int32 and pairs.is_contiguous():
        return pairs
    return _cached(("pairs", id(pairs), str(device)), (pairs,), lambda: functional._as_i32(pairs, device))


# --------------------------------------------------------------------------------------------- plan
@dc.dataclass
class Plan:
    fns: list
    model: _lib.Model
    term_mask: int
    hybrid: bool

    # the packed bank(s) when they were produced outside the configurations (mythos_b200.energy.theta_tape replays the
    # theta -> bank chain in C++); None: pack from fn.params
    bank: torch.Tensor | None = None

    def params_vector(self) -> torch.Tensor:
        """(n_banks*P,) float64 host vector, differentiable w.r.t. any tensor inside the configurations."""
        if self.bank is not None:
            return self.bank
        return bank_vector(self.fns, self.hybrid)

    def device_params(self, device, dtype) -> torch.Tensor:
        vec = self.params_vector()
        if vec.requires_grad or self.bank is not None:
            return vec.to(device=device, dtype=dtype)
        key = ("params", tuple(id(fn.params) for fn in self.fns), str(device), str(dtype))
        return _cached(key, tuple(fn.params for fn in self.fns), lambda: vec.to(device=device, dtype=dtype))

    def topology(self, n: int, device) -> functional.DeviceTopology:
        f0 = self.fns[0]
        extra: dict[str, Any] = {}
        for fn in self.fns:
            for k, v in fn.extra_topology().items():
                extra.setdefault(k, v)
        # the stacking term may carry its own nt_type (na1/tests/test_integration.py:252 does exactly that)
        stack_nt = None
        for fn in self.fns:
            if fn.TERM == 2 and "nt_type" in fn.params and fn.params.nt_type is not None:
                stack_nt = fn.params.nt_type
        keep = (f0.seq, f0.bonded_neighbors, extra.get("nt_type"), stack_nt, extra.get("is_end"))
        key = ("topo", tuple(id(k) for k in keep), str(device))
        return _cached(
            key,
            keep,
            lambda: functional.make_device_topology(
                n, f0.seq, f0.bonded_neighbors, device, nt_type=extra.get("nt_type"), nt_type_stack=stack_nt,
                is_end=extra.get("is_end"),
            ),
        )

    def pairs(self, device, topo: functional.DeviceTopology):
        """The unbonded pair source: an explicit list, or per-frame cell lists for the all-pairs sentinel."""
        if not (self.term_mask & _lib.UNBONDED_TERMS):
            return functional.StaticPairs(None)
        ub = next(fn.unbonded_neighbors for fn in self.fns if fn.TERM >= 3)
        from mythos_b200.input.topology import AllPairs

        if isinstance(ub, AllPairs):
            box = tuple(self.model.box)
            # per-frame device lists (one-pass rows layout) streamed through the energy kernels; the frame-resident
            # kernel can also find its pairs itself (src.in_kernel = True), which measured slower on B200
            in_kernel = getattr(ub, "in_kernel", False)
            # the in-kernel cell list keeps 2 bonded partners per nucleotide (like the reference's (N,2) mask)
            if in_kernel and topo.bonded.numel() and int(torch.bincount(topo.bonded.reshape(-1).long()).max()) > 2:
                in_kernel = False
            r_sr, r_db = support_cutoffs(self)
            tag = (self.model, r_sr, r_db, topo.nt_type if self.model.n_banks > 1 else None)
            return functional.CellListPairs(bonded=topo.bonded, box=box, r_cutoff=interaction_range(self), in_kernel=in_kernel, tag=tag)
        return functional.StaticPairs(device_pairs(ub, device))

    def pseq_inputs(self, n: int, device):
        """Probabilistic-sequence inputs of the stacking / hydrogen-bonding terms, or None for discrete sequences:
        ``(pmarg (N,4), same_w_stack (n_bp,2), same_w_hb (n_bp,2), bp_of, within, term bits)`` -- the three real arrays
        differentiable in ``pseq`` and in the terms' weight tables (``mythos_b200.energy.pseq``)."""
        from mythos_b200.energy import pseq as kpseq

        users = [fn for fn in self.fns if fn.TERM in (2, 4) and "pseq" in fn.params and fn.params.pseq is not None]
        if not users:
            return None
        if self.hybrid:
            raise NotImplementedError("probabilistic sequences with the hybrid (NA1) model are not on the CUDA path")
        ps, sc = users[0].params.pseq, users[0].params.pseq_constraints
        for fn in users[1:]:
            same = fn.params.pseq_constraints is sc and all(a is b or torch.equal(torch.as_tensor(a), torch.as_tensor(b))
                                                            for a, b in zip(fn.params.pseq, ps))
            if not same:
                raise NotImplementedError("stacking and hydrogen bonding must share one probabilistic sequence")
        if sc is None or sc.n_nucleotides != n:
            raise ValueError("pseq_constraints must be provided when pseq is provided." if sc is None else
                             f"pseq_constraints describe {sc.n_nucleotides} nucleotides, the body has {n}")
        pmarg = kpseq.marginals(ps, sc)
        zero = torch.zeros((sc.n_bp, 2), dtype=torch.float64)
        same_s = same_h = zero
        terms = 0
        for fn in users:
            if fn.TERM == 2:
                same_s, terms = kpseq.same_pair_weights(ps, fn.params.eps_stack, sc), terms | (1 << 2)
            else:
                same_h, terms = kpseq.same_pair_weights(ps, fn.params.eps_hb_weights, sc), terms | (1 << 4)
        bp_of, within = kpseq.index_arrays(sc)
        key = ("pseq_idx", id(sc), str(device))
        bp_dev, within_dev = _cached(key, (sc,), lambda: (functional._as_i32(bp_of, device), functional._as_i32(within, device)))
        return pmarg, same_s, same_h, bp_dev, within_dev, terms

    def evaluate(self, center: torch.Tensor, quat: torch.Tensor) -> torch.Tensor:
        """(F,8) per-term energies; differentiable in center, quat and the configurations' tensors."""
        _lib.require_cuda(center, "RigidBody.center")
        dev, dtype = center.device, center.dtype
        topo = self.topology(center.shape[1], dev)
        pq = self.pseq_inputs(center.shape[1], dev)
        if pq is not None:
            return functional.pseq_energy_terms(self.model, topo, center, quat, self.device_params(dev, dtype), self.pairs(dev, topo),
                                                self.term_mask, *pq)
        return functional.energy_terms(
            self.model, topo, center, quat, self.device_params(dev, dtype), self.pairs(dev, topo), self.term_mask
        )

    def evaluate_total(self, center: torch.Tensor, quat: torch.Tensor, weights: torch.Tensor, observables=None) -> torch.Tensor:
        """(F,) weighted total energies with the parameter-gradient rows produced in the same pass (DiffTRe shape).

        ``center`` / ``quat`` may also be PINNED HOST tensors (stored trajectory frames): they are then streamed to the
        current CUDA device chunk by chunk on a copy stream, overlapped with the kernels of the previous chunk."""
        if center.is_cuda:
            dev = center.device
        elif center.is_pinned() and quat.is_pinned() and torch.cuda.is_available():
            dev = torch.device("cuda", torch.cuda.current_device())
        else:
            _lib.require_cuda(center, "RigidBody.center (or pinned host memory, which is streamed)")
        dtype = center.dtype
        topo = self.topology(center.shape[1], dev)
        if self.pseq_inputs(center.shape[1], dev) is not None:  # probabilistic sequences: generic pair kernel, per-term route
            _lib.require_cuda(center, "RigidBody.center")
            return self.evaluate(center, quat) @ weights.to(device=dev, dtype=dtype)
        return functional.frame_energies(
            self.model, topo, center, quat, self.device_params(dev, dtype), self.pairs(dev, topo), weights, self.term_mask,
            observables,
        )


def evaluate_with_observables(plan: "Plan", center: torch.Tensor, quat: torch.Tensor, weights: torch.Tensor, observables=None) -> torch.Tensor:
    """``plan.evaluate_total`` with an optional ``ObservableSet`` evaluated in the same pass and published to its members
    (device-resident frames only: the result is remembered per frames tensor)."""
    if observables is None or not center.is_cuda:
        return plan.evaluate_total(center, quat, weights)
    req = observables.request(center.device)
    e = plan.evaluate_total(center, quat, weights, req)
    observables.publish(center, req)
    return e


def interaction_range(plan: "Plan") -> float:
    """Upper bound on the centre-of-mass distance at which any enabled unbonded term can be non-zero:
    max over terms of (site-pair cutoff + both site offsets), over all banks / flavours (unit quaternions)."""
    vec = plan.params_vector().detach()
    P = _lib.param_count()
    idx = {n: i for i, n in enumerate(_lib.param_names())}
    n_fl = 2 if plan.hybrid else 1
    off = {"back": 0.0, "base": 0.0, "stack": 0.0}
    for k in range(n_fl):
        g = plan.model.geom[k]
        off["back"] = max(off["back"], float(sum(x * x for x in g.back) ** 0.5))
        off["base"] = max(off["base"], abs(g.base))
        off["stack"] = max(off["stack"], abs(g.stack))
    reach = 0.0
    for b in range(plan.model.n_banks):
        v = vec[b * P:(b + 1) * P]

        def p(name):
            return float(v[idx[name]])

        m = plan.term_mask
        if m & (1 << 3):
            reach = max(reach, p("unbonded_excluded_volume.dr_c_backbone") + 2 * off["back"],
                        p("unbonded_excluded_volume.dr_c_base") + 2 * off["base"],
                        p("unbonded_excluded_volume.dr_c_back_base") + off["back"] + off["base"],
                        p("unbonded_excluded_volume.dr_c_base_back") + off["back"] + off["base"])
        if m & (1 << 4):
            reach = max(reach, p("hydrogen_bonding.dr_c_high_hb") + 2 * off["base"])
        if m & (1 << 5):
            reach = max(reach, p("cross_stacking.dr_c_high_cross") + 2 * off["base"])
        if m & (1 << 6):
            reach = max(reach, p("coaxial_stacking.dr_c_high_coax") + 2 * off["stack"])
        if m & (1 << 7) and plan.model.forms[b].has_debye:
            reach = max(reach, p("debye.r_cut") + 2 * off["back"])
    return reach * (1.0 + 1e-9) + 1e-9


@lru_cache(maxsize=1)
def _param_index() -> dict:
    return {n: i for i, n in enumerate(_lib.param_names())}


def support_cutoffs(plan: "Plan") -> tuple[float, float]:
    """(centre cutoff of the short-range terms, backbone-site cutoff of Debye-Hueckel), maxima over banks and flavours: the
    supports the neighbour build tags pairs with (MB_NL_TAG_SUPPORTS).  Same formula as the kernels' own short-range cutoff."""
    vec = plan.params_vector().detach().tolist()  # (one host read: this runs in front of the first launch of every pass)
    P = _lib.param_count()
    idx = _param_index()
    ob = oh = os_ = 0.0
    for k in range(2 if plan.hybrid else 1):
        g = plan.model.geom[k]
        ob = max(ob, float(sum(x * x for x in g.back) ** 0.5))
        oh, os_ = max(oh, abs(g.base)), max(os_, abs(g.stack))
    m = plan.term_mask
    r = r_db = 0.0
    for b in range(plan.model.n_banks):
        v = vec[b * P:(b + 1) * P]

        def p(name):
            return float(v[idx[name]])

        if m & (1 << 3):
            r = max(r, p("unbonded_excluded_volume.dr_c_backbone") + 2 * ob, p("unbonded_excluded_volume.dr_c_base") + 2 * oh,
                    max(p("unbonded_excluded_volume.dr_c_back_base"), p("unbonded_excluded_volume.dr_c_base_back")) + ob + oh)
        if m & (1 << 4):
            r = max(r, p("hydrogen_bonding.dr_c_high_hb") + 2 * oh)
        if m & (1 << 5):
            r = max(r, p("cross_stacking.dr_c_high_cross") + 2 * oh)
        if m & (1 << 6):
            r = max(r, p("coaxial_stacking.dr_c_high_coax") + 2 * os_)
        if (m & (1 << 7)) and plan.model.forms[b].has_debye:
            r_db = max(r_db, p("debye.r_cut") * (1.0 + 1e-9) + 1e-12)
    return r * (1.0 + 2e-6) + 1e-9, r_db


def _kw_key(v):
    # value identity of a transform_fn keyword (geometry constants are tensors: their bytes, not their repr -- formatting 40
    # tensors cost 4 ms per DiffTRe step)
    if isinstance(v, torch.Tensor):
        t = v.detach()
        return ("t", tuple(t.shape), str(t.dtype), (t if not t.is_cuda else t.cpu()).contiguous().numpy().tobytes())
    return repr(v)


def _prop_key(fn) -> tuple:
    box = space.box_of(fn.displacement_fn)
    kw = getattr(fn.transform_fn, "keywords", {}) or {}
    return (id(getattr(fn.transform_fn, "func", fn.transform_fn)), tuple((k, _kw_key(kw[k])) for k in sorted(kw)), box, id(fn.seq),
            id(fn.bonded_neighbors), id(fn.unbonded_neighbors), fn.HYBRID)


def fusable_groups(fns: list) -> list[list[int]]:
    """Partition term indices into groups that can share one fused launch (same inputs, distinct kernel terms)."""
    groups: list[tuple[tuple, set, list[int]]] = []
    for k, fn in enumerate(fns):
        key = _prop_key(fn)
        for gkey, used, members in groups:
            if gkey == key and fn.TERM not in used:
                used.add(fn.TERM)
                members.append(k)
                break
        else:
            groups.append((key, {fn.TERM}, [k]))
    return [g[2] for g in groups]


def plan_for(fns: list) -> Plan:
    """Build the launch plan of a fusable group of term objects."""
    f0 = fns[0]
    if f0.transform_fn is None:
        raise _lib.MythosB200Error("energy functions need transform_fn (the nucleotide geometry) to run")
    kind, geoms = geometry_of(f0.transform_fn)
    hybrid = any(fn.HYBRID for fn in fns)
    if hybrid != all(fn.HYBRID for fn in fns):
        raise _lib.MythosB200Error("cannot mix NA1 (hybrid) and single-model terms in one launch")
    if hybrid and kind != "na1":
        raise _lib.MythosB200Error("NA1 terms need the HybridNucleotide transform_fn")
    m = _lib.Model()
    m.n_banks = 3 if hybrid else 1
    for k, g in enumerate(geoms[:2]):
        m.geom[k] = g
    box = space.box_of(f0.displacement_fn)
    for d in range(3):
        m.box[d] = box[d]
    mask = 0
    for fn in fns:
        if fn.TERM < 0:
            raise _lib.MythosB200Error(f"{type(fn).__name__} has no kernel term")
        mask |= 1 << fn.TERM
    if hybrid:
        forms = (
            (STACK_DNA, CROSS_DNA1, COAX_DNA2),  # DNA bank: dna2 stacking / dna1 cross / dna2 coax (na1/*.py)
            (STACK_RNA, CROSS_RNA2, COAX_DNA1),  # RNA bank
            (STACK_DNA, CROSS_DNA1, COAX_DNA1),  # hybrid bank (no bonded terms)
        )
        for b, (s, c, x) in enumerate(forms):
            m.forms[b].stack_form, m.forms[b].cross_form, m.forms[b].coax_form, m.forms[b].has_debye = s, c, x, 1
        m.geom[0].use_back_stack = 1  # na1 stacking uses dna2.Stacking for the DNA bank (na1/stacking.py:203)
    else:
        form = {"stack_form": STACK_DNA, "cross_form": CROSS_DNA1, "coax_form": COAX_DNA1, "use_back_stack": 0}
        for fn in fns:
            form.update(fn.FORM)
        m.forms[0].stack_form = form["stack_form"]
        m.forms[0].cross_form = form["cross_form"]
        m.forms[0].coax_form = form["coax_form"]
        m.forms[0].has_debye = 1 if mask & (1 << 7) else 0
        m.geom[0].use_back_stack = 1 if (form["use_back_stack"] and kind == "dna2") else 0
    hce = 0
    for fn in fns:
        if fn.TERM == 7:
            hce = 1 if bool(fn.params.half_charged_ends) else 0
    m.half_charged_ends = hce
    return Plan(fns=list(fns), model=m, term_mask=mask, hybrid=hybrid)
