"""Differentiable entry to the fused CUDA energy kernels (the role ``jax.ffi.ffi_call`` + ``custom_vjp`` plays
in the XLA integration, here with torch.autograd for device memory and the chain rule).

``energy_terms(...)`` returns the per-term energies ``(F, 8)`` of ``F`` frames in one launch pair
(bonded + unbonded).  Its backward is remat-style: the forward saves only its inputs, the backward is a second
C-ABI call that recomputes the pairs and writes ``d/dcenter (F,N,3)``, ``d/dquat (F,N,4)`` and
``d/dparams (n_banks*P,)`` for the incoming cotangent ``(F, 8)`` -- exactly the differentiation contract of
SURVEY 8b.  There is no other implementation behind this function: without the shared object or without a CUDA
device it raises.
"""

from __future__ import annotations

import ctypes as C
import collections
import dataclasses as dc
import os
import threading

import torch

from mythos_b200 import _lib


@dc.dataclass(frozen=True)
class DeviceTopology:
    """Integer inputs of the kernels, resident on the device (int32, contiguous)."""

    n: int
    seq: torch.Tensor
    bonded: torch.Tensor  # (B,2)
    nt_type: torch.Tensor | None = None
    nt_type_stack: torch.Tensor | None = None
    is_end: torch.Tensor | None = None


def _as_i32(x, device) -> torch.Tensor | None:
    if x is None:
        return None
    t = torch.as_tensor(x)
    return t.to(device=device, dtype=torch.int32).contiguous()


def make_device_topology(n, seq, bonded, device, nt_type=None, nt_type_stack=None, is_end=None) -> DeviceTopology:
    b = _as_i32(bonded, device)
    b = b.reshape(-1, 2) if b is not None else torch.zeros((0, 2), dtype=torch.int32, device=device)
    return DeviceTopology(
        n=int(n),
        seq=_as_i32(seq, device),
        bonded=b,
        nt_type=_as_i32(nt_type, device),
        nt_type_stack=_as_i32(nt_type_stack, device),
        is_end=_as_i32(is_end, device),
    )


def _launch(
    model: _lib.Model,
    topo: DeviceTopology,
    center: torch.Tensor,
    quat: torch.Tensor,
    params: torch.Tensor,
    pairs: torch.Tensor | None,
    pair_frame_stride: int,
    term_mask: int,
    cot: torch.Tensor | None,
    want_terms: bool,
    want_pos_grad: bool,
    want_param_grad: bool,
    per_frame_param_grad: bool = False,
    pair_count: torch.Tensor | None = None,
    flags: int = 0,
    all_pairs_cutoff: float = 0.0,
    out: tuple | None = None,
    pair_split: torch.Tensor | None = None,
    observables=None,
    pseq=None,
):
    _lib.require_cuda(center, "center")
    F, N = center.shape[0], center.shape[1]
    dtype, dev = center.dtype, center.device
    sfx = _lib.suffix(dtype)
    if N != topo.n:
        raise _lib.MythosB200Error(f"body has {N} nucleotides, topology has {topo.n}")
    center = center.contiguous()
    quat = quat.to(dtype).contiguous()
    params = params.to(device=dev, dtype=dtype).contiguous()
    np_ = model.n_banks * _lib.param_count()
    if params.numel() != np_:
        raise _lib.MythosB200Error(f"params has {params.numel()} entries, expected {np_}")
    terms = torch.empty((F, _lib.N_TERMS), dtype=dtype, device=dev) if want_terms else None
    if out is not None:  # caller-owned gradient buffers (MD loop): (d_center, d_quat), accumulated into
        d_center, d_quat = out
    else:
        d_center = torch.empty_like(center) if want_pos_grad else None
        d_quat = torch.empty_like(quat) if want_pos_grad else None
    d_params = None
    stride = 0
    if want_param_grad:
        if per_frame_param_grad:
            d_params = torch.empty((F, np_), dtype=dtype, device=dev)
            stride = np_
        else:
            d_params = torch.empty((np_,), dtype=dtype, device=dev)
    if cot is not None:
        cot = cot.to(device=dev, dtype=dtype).contiguous()
    cap = 0
    if pairs is not None and pairs.numel() > 0:
        cap = pairs.shape[-1]
    a = _lib.EnergyArgs()
    a.model = C.pointer(model)
    a.n, a.n_frames = N, F
    a.center, a.quat = center.data_ptr(), quat.data_ptr()
    a.seq = topo.seq.data_ptr()
    a.nt_type = _lib.ptr(topo.nt_type)
    a.nt_type_stack = _lib.ptr(topo.nt_type_stack)
    a.is_end = _lib.ptr(topo.is_end)
    a.bonded = topo.bonded.data_ptr() if topo.bonded.numel() else None
    a.n_bonded = topo.bonded.shape[0]
    a.pairs = pairs.data_ptr() if cap else None
    a.pair_capacity = cap
    a.pair_frame_stride = pair_frame_stride
    a.params = params.data_ptr()
    a.cot = _lib.ptr(cot)
    a.term_mask = term_mask
    a.flags = flags
    a.pair_count = _lib.ptr(pair_count) if cap else None
    a.all_pairs_cutoff = float(all_pairs_cutoff)
    a.terms = _lib.ptr(terms)
    a.d_center = _lib.ptr(d_center)
    a.d_quat = _lib.ptr(d_quat)
    a.d_params = _lib.ptr(d_params)
    a.d_params_frame_stride = stride
    ws = None
    a.pair_split = _lib.ptr(pair_split) if cap else None
    pseq_struct = None
    if pseq is not None:  # PseqInputs: device tensors of the probabilistic-sequence weights (+ gradient buffers, accumulated into)
        pseq_struct = pseq.struct(dtype)
        a.pseq = C.addressof(pseq_struct)
        flags |= _lib.FLAG_GENERIC_KERNEL
        a.flags = flags
    obs_spec = None
    if observables is not None:  # (ObservableRequest, out (F,4)): fused epilogue of the frame-resident kernel, else a launch behind it
        obs_spec = observables[0].struct()
        a.observables = C.addressof(obs_spec)
        a.observables_out = observables[1].data_ptr()
    esize = 8 if dtype == torch.float64 else 4
    tagged_list = bool(flags & _lib.FLAG_TAGGED_PAIRS)
    # will the frame-resident kernel take this call?  (same rule as frame_kernel_eligible in csrc/frame_kernels.cu)  It needs
    # only its own scratch (nothing in the default build); lending it the LIST kernels' workspace -- 0.66 GB per 1024 frames,
    # a fresh block per launch -- made the caching allocator cudaMalloc at unpredictable passes (~100 ms each).
    frame_route = (not want_pos_grad and model.n_banks == 1 and not (flags & (_lib.FLAG_GENERIC_KERNEL | _lib.FLAG_LIST_KERNEL))
                   and N <= 60000 and (not tagged_list or N < 16384)
                   and bool(_lib.lib().mythos_b200_frame_kernel_fits(N, esize, 1 if want_param_grad else 0)))
    if frame_route:
        need = int(_lib.lib().mythos_b200_energy_workspace_bytes(N, F, 0, esize)) if want_param_grad else 0
        if need:
            ws = torch.empty(need, dtype=torch.uint8, device=dev)
            a.workspace, a.workspace_bytes = ws.data_ptr(), need
    elif cap and (cap * F >= 65536 or (flags & (_lib.FLAG_LIST_KERNEL | _lib.FLAG_TAGGED_PAIRS))) and not (flags & _lib.FLAG_GENERIC_KERNEL):
        # scratch of the phase-queued list kernels (caller-owned, as everywhere in the C-ABI); the caching allocator makes
        # this a pointer bump, and it is graph-capture safe
        need = int(_lib.lib().mythos_b200_energy_workspace_bytes(N, F, cap, esize))
        ws = torch.empty(max(need, 16), dtype=torch.uint8, device=dev)
        a.workspace, a.workspace_bytes = ws.data_ptr(), need
    fn = getattr(_lib.lib(), f"mythos_b200_energy_{sfx}")
    with torch.cuda.device(dev):
        _lib.check(fn(_lib.current_stream(dev), C.byref(a)), "mythos_b200_energy")
    return terms, d_center, d_quat, d_params


PACKED_SLOTS = True  # False: padded warp slots also on the frame-resident route (A/B tests)


@dc.dataclass(frozen=True)
class StaticPairs:
    """An explicit unbonded list: (2,U) shared by all frames, or (F,2,U) one list per frame."""

    pairs: torch.Tensor | None

    def __post_init__(self):
        if self.pairs is not None and self.pairs.numel() > 0 and (self.pairs.dim() not in (2, 3) or self.pairs.shape[-2] != 2):
            raise _lib.MythosB200Error(f"pair list must have shape (2,U) or (F,2,U), got {tuple(self.pairs.shape)} "
                                       "(a (U,2) `topology.unbonded_neighbors` must be transposed first)")

    def chunk(self, sl: slice, center: torch.Tensor):
        if self.pairs is None or self.pairs.numel() == 0:
            return None, 0, None
        if self.pairs.dim() == 3:
            return self.pairs[sl], 2 * self.pairs.shape[-1], None
        return self.pairs, 0, None


class _SizingMemo:
    """What the pair sources learn about a system -- list capacities, warp-slot geometry, whether float32 tagged builds are
    safe -- remembered per (device, system size, cutoffs, box).  Energy functions are rebuilt by every ``with_params()``
    and each rebuild makes a fresh ``CellListPairs``: without this memo a source that had to grow its capacity (a later
    frame holds more pairs than frame 0) would be discarded together with what it learnt, and the next pass would
    overflow on the same frame again, forever.  Lock-guarded: XLA / user threads may evaluate concurrently."""

    def __init__(self):
        import threading

        self._lock = threading.Lock()
        self._data: dict = {}

    def get(self, key) -> dict:
        with self._lock:
            return dict(self._data.get(key, ()))

    def update(self, key, **values) -> None:
        with self._lock:
            if len(self._data) > 256 and key not in self._data:
                self._data.pop(next(iter(self._data)))
            self._data.setdefault(key, {}).update(values)

    def clear(self) -> None:
        with self._lock:
            self._data.clear()


_SIZING = _SizingMemo()
_PAIR_SCRATCH: dict = {}


_NL_SCRATCH: dict = {}


def _scratch_nl_workspace(device, n: int, n_frames: int) -> torch.Tensor:
    """The neighbour build's scratch, one reused buffer per (device, stream) (pair sources are rebuilt every pass)."""
    key = (str(device), torch.cuda.current_stream(device).cuda_stream)
    need = int(_lib.lib().mythos_b200_nl_workspace_bytes(n, n_frames))
    buf = _NL_SCRATCH.get(key)
    if buf is None or buf.numel() < need:
        _NL_SCRATCH[key] = buf = torch.empty(int(need * 1.1) + 1024, dtype=torch.uint8, device=device)
    return buf


def _scratch_pairs(device, n_frames: int, cap: int) -> torch.Tensor:
    """(F,2,cap) int32 view of the per-(device, stream) scratch buffer for pair lists that live for one launch."""
    key = (str(device), torch.cuda.current_stream(device).cuda_stream)
    need = n_frames * 2 * cap
    buf = _PAIR_SCRATCH.get(key)
    if buf is None or buf.numel() < need:
        _PAIR_SCRATCH[key] = buf = torch.empty(int(need * 1.1) + 1024, dtype=torch.int32, device=device)
    return buf[:need].view(n_frames, 2, cap)
MAX_PASS_REPEATS = 6  # a pass is repeated when a pair list overflowed; capacities grow geometrically, so this is generous


@dc.dataclass
class CellListPairs:
    """All-pairs semantics (the reference's ``topology.unbonded_neighbors``) realised per frame on the device.

    Every unbonded term has compact support, so evaluating the pairs inside ``r_cutoff`` (the interaction range of
    the current parameters, ``mythos_b200.energy.model.interaction_range``) gives the same energies as the
    reference's N(N-1)/2 list.  The list is rebuilt per frame by the cell-list kernels and never leaves the GPU.
    """

    bonded: torch.Tensor
    box: tuple[float, float, float]
    r_cutoff: float
    capacity: int = 0  # 0 = size from the first frame
    workspace: torch.Tensor | None = None
    max_count: int = 0
    in_kernel: bool = False  # True: let the frame-resident kernel find the pairs itself (in-kernel cell list) where it applies
    # support tagging (MB_NL_TAG_SUPPORTS): (model, short-range centre cutoff, Debye backbone-site cutoff) or None.  With it
    # the build keeps only pairs inside the support of some term and tags which; the frame-resident kernel then queues
    # them without touching coordinates.  Used when that kernel applies (one bank, no position gradients).
    tag: tuple | None = None
    last_valid_count: torch.Tensor | None = None
    slot_geometry: tuple | None = None  # warp-slot builds: ((lane_slots, slot_width) of the short-range build, of the Debye build)
    _slot_stats: list = dc.field(default_factory=list)
    tag_float32: bool = True  # run the (superset) tagged builds in float32; cleared by verify() if the system is too extended
    _extents: list = dc.field(default_factory=list)
    tag_for_list_kernels: bool = False  # opt-in: support-tagged lists also on the list-kernel route (forces, large systems)
    tagged_capacity: int = 0
    last_split: torch.Tensor | None = None  # (F) of the last tagged chunk: entries before it are short-range pairs
    _last_tagged: bool = False
    _pending: list = dc.field(default_factory=list)
    _queued: tuple | None = None  # read-back of a pass's flags in flight: (pinned vector, event, ...), see enqueue_verification

    _memo_key: tuple | None = None
    keep_lists: bool = False  # the lists of this pass will be remembered: they must not share the reused scratch buffer
    _cache_candidates: list = dc.field(default_factory=list)  # lists built this pass, remembered once verify() passes

    def _settle_cache(self, good: bool) -> None:
        cands, self._cache_candidates = self._cache_candidates, []
        if good:
            for args in cands:
                _PAIR_LISTS.commit(*args)

    def _load_memo(self, device, n: int) -> None:
        """Seed this (fresh) source from what earlier sources of the same system learnt."""
        if self._memo_key is not None:
            return
        tag = None if self.tag is None else (round(float(self.tag[1]), 5), round(float(self.tag[2]), 5))
        self._memo_key = (str(device), int(n), round(float(self.r_cutoff), 5), tag, tuple(float(b) for b in self.box))
        m = _SIZING.get(self._memo_key)
        if self.capacity <= 0:
            self.capacity = int(m.get("capacity", 0))
        if self.tagged_capacity <= 0:
            self.tagged_capacity = int(m.get("tagged_capacity", 0))
        if self.slot_geometry is None:
            self.slot_geometry = m.get("slot_geometry")
        if m.get("tag_float32") is False:
            self.tag_float32 = False

    def _store_memo(self) -> None:
        if self._memo_key is not None:
            _SIZING.update(self._memo_key, capacity=self.capacity, tagged_capacity=self.tagged_capacity,
                           slot_geometry=self.slot_geometry, tag_float32=self.tag_float32)

    def _slot_chunk(self, c, sites, r_sr, r_db):
        """The two tagged builds in the one-pass warp-slot layout (``MB_NL_WARP_SLOTS``): every warp of 32 cell-ordered
        nucleotides writes its pairs into its own fixed-width slot of the list and pads the rest with N.  No count pass, no
        scan, no second walk; the frame-resident kernel skips the padding entries like any padded OrderedSparse list."""
        from mythos_b200.utils import neighbors

        n = c.shape[1]
        wpf = (n + 31) // 32

        def run(cc, ss, geo):
            (ka, wa), (kb, wb) = geo
            cap = wpf * (wa + (wb if ss is not None else 0))
            F = cc.shape[0]
            # lists that are consumed by the very next launch and then dropped live in one reused per-device buffer (stream
            # order makes that safe); a fresh 0.6 GB tensor per chunk made the caching allocator cudaMalloc a new block at
            # unpredictable passes (~100 ms each).  Lists that will be remembered (_PairListCache) own their memory.
            pairs = (torch.empty((F, 2, cap), dtype=torch.int32, device=cc.device) if self.keep_lists
                     else _scratch_pairs(cc.device, F, cap))
            count = torch.empty((F,), dtype=torch.int32, device=cc.device)
            overflow = torch.zeros((1,), dtype=torch.int32, device=cc.device)
            mra = torch.empty((F, 2), dtype=torch.int32, device=cc.device)
            mrb = torch.empty((F, 2), dtype=torch.int32, device=cc.device) if ss is not None else None
            self.workspace = _scratch_nl_workspace(cc.device, cc.shape[1], F)
            # frame-resident builds (free space, a frame's cell table and records in shared memory) pack the warps' slots back
            # to back: the consumer reads count[f] entries instead of the padded capacity (a third fewer producer steps)
            packed = PACKED_SLOTS and not any(self.box)
            if packed:
                with torch.cuda.device(cc.device):
                    packed = bool(_lib.lib().mythos_b200_nl_conditional_supported(n, max(ka, kb if ss is not None else 0), cc.element_size()))
            _, _, _, self.workspace = neighbors.build_pairs(cc, self.bonded, self.box, max(r_sr, 1e-6), 0.0, cap, self.workspace,
                                                            tag_bits=1 << 30, out=(pairs, count, overflow), max_row=mra,
                                                            warp_slots=(ka, 0, wa), packed_slots=packed)
            if ss is not None:
                # (same workspace, same bonded list as the build just enqueued: its exclusion table is reused)
                neighbors.build_pairs(ss, self.bonded, self.box, r_db, 0.0, cap, self.workspace, tag_bits=1 << 29,
                                      out=(pairs, count, overflow), max_row=mrb, warp_slots=(kb, wpf * wa, wb), reuse_exclusions=True,
                                      packed_slots=packed)
            self.last_valid_count = count  # (F) pairs actually written
            return pairs, cap, overflow, mra, mrb, (count if packed else None)

        def sized(lane_max, warp_max):
            return min(max(int(lane_max * 1.5) + 4, 8), 256), (max(int(warp_max * 1.35) + 16, 32) + 3) // 4 * 4

        # slot sizes are remembered per system (_SizingMemo): a fresh pair source must not pay a probe -- or worse, an
        # overflow and a repeated pass -- on every step
        if self.slot_geometry is None:  # probe the first frame with generous slots, size from what it needed
            _, _, _, mra, mrb, _ = run(c[:1], None if sites is None else sites[:1], ((256, 32 * 128), (256, 32 * 128)))
            a_l, a_w = (int(v) for v in mra.max(0).values.tolist())
            b_l, b_w = (int(v) for v in mrb.max(0).values.tolist()) if mrb is not None else (0, 0)
            self.slot_geometry = (sized(a_l, a_w), sized(b_l, b_w))
            self._store_memo()
        pairs, cap, overflow, mra, mrb, count = run(c, sites, self.slot_geometry)
        self._pending.append((torch.zeros((1,), dtype=torch.int32, device=c.device), overflow))
        self._slot_stats.append((mra, mrb))
        return pairs, 2 * cap, count  # (count: packed slots -- valid entries at the head; None: padded slots, scan the capacity)

    def chunk(self, sl: slice, center: torch.Tensor, quat: torch.Tensor | None = None, tagged: bool = False, slots: bool = False):
        """Enqueue the build for one chunk of frames; overflow is checked once per pass by ``verify`` (one host sync).

        ``tagged``: two builds into ONE list -- the centres at the short-range cutoff (tag bit 30) and the backbone sites
        at the Debye-Hueckel cutoff (tag bit 29, appended) -- instead of one build at the full interaction range: only
        pairs inside the support of some term are written, each labelled with the phase queue it belongs to."""
        from mythos_b200.utils import neighbors

        c = center.detach()
        tagged = bool(tagged and self.tag is not None and quat is not None)
        self._last_tagged = tagged
        self._load_memo(c.device, c.shape[1])
        if not tagged:
            if self.capacity <= 0:
                _, count, _, self.workspace = neighbors.build_pairs(c[:1], self.bonded, self.box, self.r_cutoff, 0.0, 1, self.workspace)
                self.capacity = (max(int(int(count.max().item()) * 1.06) + 64, 64) + 3) // 4 * 4  # multiple of 4: 128-bit pair stores
                self._store_memo()
            pairs, count, overflow, self.workspace = neighbors.build_pairs(
                c, self.bonded, self.box, self.r_cutoff, 0.0, self.capacity, self.workspace
            )
            self._pending.append((count, overflow))
            return pairs, 2 * self.capacity, count
        model, r_sr, r_db = self.tag[:3]
        nt_type = self.tag[3] if len(self.tag) > 3 else None
        # The tagged builds only have to be SUPERSETS of the supports (the kernels apply the exact tests again), so they run
        # in float32 with the cutoffs widened by 1e-3: half the record bytes and FP32 instead of FP64 distance arithmetic in
        # the walks.  One kernel writes both point sets as float32 brought near the origin (periodic: primary image; free
        # space: relative to the frame's first nucleotide -- distances do not change) and their largest magnitude, which
        # verify() checks: beyond 1500 length units float32 could not resolve 1e-3 and the pass is redone in float64.
        extent = None
        if self.tag_float32:
            extent = torch.zeros(1, dtype=torch.float32, device=c.device)
            c32, s32 = torch.empty(c.shape, dtype=torch.float32, device=c.device), torch.empty(c.shape, dtype=torch.float32, device=c.device)
            fn = getattr(_lib.lib(), f"mythos_b200_support_points_{_lib.suffix(c.dtype)}")
            with torch.cuda.device(c.device):
                _lib.check(fn(_lib.current_stream(c.device), C.pointer(model), c.shape[1], c.shape[0], c.contiguous().data_ptr(),
                              quat.detach().to(c.dtype).contiguous().data_ptr(), _lib.ptr(nt_type), c32.data_ptr(), s32.data_ptr(),
                              extent.data_ptr()), "mythos_b200_support_points")
            c, sites = c32, (s32 if r_db > 0 else None)
            r_sr, r_db = r_sr + 1e-3, (r_db + 1e-3 if r_db > 0 else r_db)
        else:
            sites = backbone_sites(model, c, quat.detach(), nt_type) if r_db > 0 else None

        if slots:
            if extent is not None:
                self._extents.append(extent)
            return self._slot_chunk(c, sites, r_sr, r_db)

        def both(cc, ss, cap):
            pairs, split, overflow, self.workspace = neighbors.build_pairs(cc, self.bonded, self.box, max(r_sr, 1e-6), 0.0, cap,
                                                                           self.workspace, tag_bits=1 << 30)
            count = split
            if ss is not None:
                count = torch.empty_like(split)
                neighbors.build_pairs(ss, self.bonded, self.box, r_db, 0.0, cap, self.workspace, tag_bits=1 << 29,
                                      out=(pairs, count, overflow), append_count=split)
            return pairs, count, overflow, split

        if self.tagged_capacity <= 0:
            _, count, _, _ = both(c[:1], None if sites is None else sites[:1], 4)
            self.tagged_capacity = (max(int(int(count.max().item()) * 1.06) + 64, 64) + 3) // 4 * 4
            self._store_memo()
        pairs, count, overflow, self.last_split = both(c, sites, self.tagged_capacity)
        self._pending.append((count, overflow))
        if extent is not None:
            self._extents.append(extent)
        return pairs, 2 * self.tagged_capacity, count

    def verify(self) -> bool:
        """True if every list built since the last call fitted its capacity; otherwise grows the capacity."""
        good = self._verify()
        self._settle_cache(good)
        return good


    def enqueue_verification(self) -> None:
        """Launch the reductions of everything this pass recorded (extent, longest list, overflow flags, slot statistics) and
        their copy into pinned host memory NOW, without waiting: ``verify`` later only waits for the event.  Called by
        ``deferred_verification.enqueue`` right after the forward launches, so the read-back rides behind the caller's
        own device -> host copy instead of adding a second round of tiny kernels and a second sync at the end of the step."""
        if not self._pending or self._queued is not None:
            return
        dev = self._pending[0][1].device
        parts = [torch.stack(self._extents).max().double().reshape(1) if self._extents else torch.zeros(1, dtype=torch.float64, device=dev),
                 torch.stack([c.max() for c, _ in self._pending]).max().double().reshape(1),
                 torch.stack([o[0] for _, o in self._pending]).max().double().reshape(1)]
        stats, self._slot_stats = self._slot_stats, []
        if stats:
            parts.append(torch.stack([m.max(0).values for m, _ in stats]).max(0).values.double())
            if stats[0][1] is not None:
                parts.append(torch.stack([m.max(0).values for _, m in stats]).max(0).values.double())
        dev_vec = torch.cat(parts)
        host_vec = torch.empty(dev_vec.shape, dtype=dev_vec.dtype, pin_memory=True)
        host_vec.copy_(dev_vec, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev))
        self._queued = (host_vec, ev, bool(self._extents), stats, dev_vec)
        self._extents.clear()
        self._pending.clear()

    def _verify(self) -> bool:
        self.enqueue_verification()
        if self._queued is None:
            return True
        host_vec, ev, had_extents, stats, _keep = self._queued
        self._queued = None
        ev.synchronize()
        host = host_vec.tolist()
        if had_extents and host[0] > 1500.0:  # float32 cannot resolve the 1e-3 margin out there: redo the pass with float64 builds
            self.tag_float32 = False
            self._store_memo()
            return False
        worst, flags = int(host[1]), int(host[2])
        if stats:  # warp-slot builds: bit 0 = a slot was too narrow, bit 2 = a lane row was too short
            if flags & 2:
                raise _lib.MythosB200Error("neighbour build: a nucleotide has more than 4 bonded partners")
            a = host[3:5]
            b = host[5:7] if len(host) > 5 else [0, 0]
            (ka, wa), (kb, wb) = self.slot_geometry

            def fit(lane, warp, k, w, grow_only):
                # wanted sizes: a comfortable margin over what this whole pass needed; slots are also tightened when they
                # turn out much wider than that (padding entries cost the consumer a queue step per 3072 of them)
                k2, w2 = min(int(lane * 1.5) + 4, 256), (int(warp * 1.12) + 16 + 3) // 4 * 4
                if grow_only:
                    return max(k, k2), max(w, w2)
                return max(k, k2), (w2 if w > 1.1 * w2 else max(w, w2))

            over = bool(flags & 5)
            self.slot_geometry = (fit(a[0], a[1], ka, wa, over), fit(b[0], b[1], kb, wb, over))
            self._store_memo()
            return not over
        if flags & 2:
            raise _lib.MythosB200Error("neighbour build: a nucleotide has more than 4 bonded partners")
        self.max_count = max(self.max_count, worst)
        cap_attr = "tagged_capacity" if self._last_tagged else "capacity"
        if worst <= getattr(self, cap_attr):
            return True
        setattr(self, cap_attr, (int(worst * 1.06) + 64 + 3) // 4 * 4)  # jax_md would report did_buffer_overflow; here the pass re-runs
        self._store_memo()
        return False


def backbone_sites(model, center: torch.Tensor, quat: torch.Tensor, nt_type: torch.Tensor | None = None) -> torch.Tensor:
    """(F,N,3) backbone interaction sites of (center, quat) (one fused kernel); ``nt_type`` (N) int32 picks the flavour
    per nucleotide for the three-bank model."""
    out = torch.empty_like(center)
    fn = getattr(_lib.lib(), f"mythos_b200_backbone_sites_{_lib.suffix(center.dtype)}")
    with torch.cuda.device(center.device):
        _lib.check(fn(_lib.current_stream(center.device), C.pointer(model), center.shape[0] * center.shape[1],
                      center.contiguous().data_ptr(), quat.to(center.dtype).contiguous().data_ptr(), out.data_ptr(),
                      _lib.ptr(nt_type), center.shape[1]), "mythos_b200_backbone_sites")
    return out


class _PairListCache:
    """Per-frame pair lists of stored trajectory frames, kept between passes.

    DiffTRe evaluates the SAME stored frames again and again -- reference energies, new energies and the loss pass of
    every optimiser step until the trajectory is resampled (``mythos/optimization/objective.py:298-303,345-361``) -- and
    a pair list depends on the parameters only through the cutoffs it was built with.  Lists are therefore built with
    the cutoffs widened by ``PAIR_LIST_MARGIN`` and remembered per (frames tensor, chunk): a later pass over the same
    tensor objects whose cutoffs still fit inside the built ones streams the remembered lists and launches no
    neighbour kernels at all (the energy kernels re-apply every term's exact support, so a superset list changes nothing
    -- tests compare cached and rebuilt passes).  Keyed by the identity of the ``center`` / ``quat`` tensor objects
    (weak references: an entry dies with its frames) and their version counters (an in-place update invalidates it);
    bounded by ``PAIR_LIST_CACHE_GB``.  Entries are committed only after the pass that built them verified its lists."""

    def __init__(self):
        import threading

        self._lock = threading.Lock()
        self._entries: dict = {}  # id(center) -> dict(ref_c, ref_q, version, sig, chunks={(lo,hi): value}, bytes)

    def _alive(self, e, center, quat) -> bool:
        return (e is not None and e["ref_c"]() is center and e["ref_q"]() is quat
                and e["version"] == (center._version, quat._version))

    def lookup(self, center, quat, sig, cutoffs, lo, hi):
        """The remembered lists of frames [lo, hi) if they were built for the same system with cutoffs >= `cutoffs`."""
        with self._lock:
            e = self._entries.get(id(center))
            if not self._alive(e, center, quat):
                self._entries.pop(id(center), None)
                return None
            if e["sig"] != sig or any(need > have for need, have in zip(cutoffs, e["cutoffs"])):
                return None
            return e["chunks"].get((lo, hi))

    def built_cutoffs(self, center, quat, sig):
        with self._lock:
            e = self._entries.get(id(center))
            return e["cutoffs"] if self._alive(e, center, quat) and e["sig"] == sig else None

    def commit(self, center, quat, sig, cutoffs, lo, hi, value, nbytes) -> None:
        import weakref

        budget = PAIR_LIST_CACHE_GB * 2**30
        with self._lock:
            e = self._entries.get(id(center))
            if not self._alive(e, center, quat) or e["sig"] != sig or e["cutoffs"] != cutoffs:
                key = id(center)
                e = {"ref_c": weakref.ref(center, lambda _r, k=key: self._drop(k)), "ref_q": weakref.ref(quat),
                     "version": (center._version, quat._version), "sig": sig, "cutoffs": cutoffs, "chunks": {}, "bytes": 0}
                self._entries[key] = e
            total = sum(x["bytes"] for x in self._entries.values())
            while total + nbytes > budget and len(self._entries) > 1:  # oldest trajectories go first
                k = next(k for k in self._entries if k != id(center))
                total -= self._entries.pop(k)["bytes"]
            if total + nbytes > budget:
                return
            e["chunks"][(lo, hi)] = value
            e["bytes"] += nbytes

    def _drop(self, key) -> None:
        with self._lock:
            self._entries.pop(key, None)

    def clear(self) -> None:
        with self._lock:
            self._entries.clear()

    def nbytes(self) -> int:
        with self._lock:
            return sum(x["bytes"] for x in self._entries.values())


PAIR_LIST_CACHE_GB = float(os.environ.get("MYTHOS_B200_PAIR_LIST_CACHE_GB", "24"))  # 0 disables the cache
PAIR_LIST_MARGIN = 0.01  # cutoffs of cached lists are widened by this fraction, so small parameter updates keep them valid
_PAIR_LISTS = _PairListCache()


# frames per launch group: 14 waves of one CTA per SM for frames that are already on the device (fewer launch tails: a
# launch ends when its slowest CTA does; measured 23.4 -> 22.8 ms per 8192 frames against 8 waves), 8 waves for frames
# streamed from the host (larger chunks measured slower end to end).  Bounds the pair-list buffer (~1 GB at N=2k).
FRAME_CHUNK = int(os.environ.get("MYTHOS_B200_FRAME_CHUNK", "2072"))
STREAM_CHUNK = int(os.environ.get("MYTHOS_B200_STREAM_CHUNK", "1184"))


STREAM_FIRST_CHUNK = 148  # frames of the first chunk of a pass over pinned HOST frames (one wave of one CTA per SM)
STREAM_RAMP = (1, 2, 3, 5, 7, 10, 14)  # chunk sizes of a streamed pass in units of the first chunk, then the regular size
STREAM_AHEAD_BYTES = int(float(os.environ.get("MYTHOS_B200_STREAM_AHEAD_GB", "2")) * 2**30)  # host -> device copies queued ahead of the kernels


def _chunks(n_frames: int, source, streamed: bool = False) -> list[slice]:
    """Frame ranges of one pass.  Frames streamed from pinned host memory start with a one-wave chunk and grow by about
    1.4x per chunk up to the regular size: the copies of ALL chunks are queued on the copy stream ahead of the kernels
    (``STREAM_AHEAD_BYTES``), a frame's copy takes ~0.75 of its evaluation, so a chunk that is at most ~1.33x the frames
    evaluated before it has arrived by the time the kernels get to it (doubling, the earlier ramp, left the kernels waiting
    for the copy at every step of the ramp) -- also when a rank's whole block is smaller than one regular chunk."""
    step = FRAME_CHUNK if (isinstance(source, CellListPairs) or source is CellListPairs) else 65535
    if not streamed:
        return [slice(lo, min(lo + step, n_frames)) for lo in range(0, n_frames, step)]
    step = min(step, STREAM_CHUNK)
    first = min(STREAM_FIRST_CHUNK, step)
    sizes = [min(w * first, step) for w in STREAM_RAMP if w * first < step] + [step]
    out, lo, k = [], 0, 0
    while lo < n_frames:
        size = sizes[min(k, len(sizes) - 1)]
        hi = min(lo + size, n_frames)
        if n_frames - hi < size // 2:  # do not leave a sliver for a chunk of its own
            hi = n_frames if n_frames - lo <= step else hi
        out.append(slice(lo, hi))
        lo, k = hi, k + 1
    return out


def _run(model, topo, center, quat, params, source, term_mask, cot, want_terms, want_pos, want_par, per_frame_par, flags=0,
         observables=None, pseq=None):
    """Chunked launch over frames; concatenates / sums the per-chunk outputs.  ``observables``: an ``ObservableRequest``
    whose ``out`` receives the (F,4) per-frame observables evaluated in the same pass."""
    if observables is not None:
        observables.out = torch.empty((center.shape[0], _lib.N_OBS), dtype=center.dtype, device=params.device)
    if pseq is not None:
        flags |= _lib.FLAG_GENERIC_KERNEL  # probabilistic sequence weights live in the generic pair kernel
    if isinstance(source, CellListPairs) and not want_pos and model.n_banks == 1 and source.in_kernel and not (flags & _lib.FLAG_GENERIC_KERNEL):
        # all-pairs mode inside the frame-resident kernel: the CTA finds its own pairs, no list in HBM
        try:
            outs = [
                _launch(model, topo, center[sl], quat[sl], params, None, 0, term_mask, None if cot is None else cot[sl],
                        want_terms, False, want_par, per_frame_par, None, flags, source.r_cutoff,
                        observables=None if observables is None else (observables, observables.out[sl]))
                for sl in _chunks(center.shape[0], None)
            ]
            return _merge(outs, want_terms, False, want_par, per_frame_par)
        except _lib.MythosB200Error as err:
            if getattr(err, "status", None) != 3:  # MB_ECAPACITY: frame too large for shared memory -> device lists
                raise
            source.in_kernel = False
    # support-tagged lists: for the frame-resident kernel (one bank, no position gradients, small frames); the list
    # kernels of large systems take them too (any bank count, forces), but for ONE configuration the second build's fixed
    # launch cost outweighs the pairs it saves (measured: 0.92 vs 0.85 ms at 100k nucleotides), so that is opt-in
    tagged = (isinstance(source, CellListPairs) and source.tag is not None
              and not (flags & (_lib.FLAG_GENERIC_KERNEL | _lib.FLAG_LIST_KERNEL))
              and ((not want_pos and model.n_banks == 1 and center.shape[1] < 16384
                    and _lib.lib().mythos_b200_frame_kernel_fits(center.shape[1], center.element_size(), 1 if want_par else 0))
                   or source.tag_for_list_kernels))
    frame_route = tagged and not want_pos and model.n_banks == 1 and center.shape[1] < 16384 and not source.tag_for_list_kernels
    # frames in pinned host memory are streamed: chunk k+1 is copied on a side stream while chunk k is evaluated
    streamed = not center.is_cuda
    if streamed and want_pos:
        raise _lib.MythosB200Error("position gradients need device-resident frames")
    dev = params.device
    # stored frames that are evaluated again and again (DiffTRe) keep their pair lists between passes (_PairListCache)
    cacheable = (isinstance(source, CellListPairs) and PAIR_LIST_CACHE_GB > 0 and center.dim() == 3 and center.shape[0] > 1
                 and not center.requires_grad and not quat.requires_grad)
    # (today's cutoffs, read once: a repeated pass must not widen the already widened ones again)
    need_tagged = (float(source.tag[1]), float(source.tag[2])) if cacheable and source.tag is not None else None
    need_plain = (float(source.r_cutoff),) if cacheable else None
    repeats = 0
    while True:
        outs = []
        repeats += 1
        if repeats > MAX_PASS_REPEATS:
            raise _lib.MythosB200Error(f"pair lists still overflow after {MAX_PASS_REPEATS} passes "
                                       f"(capacity {getattr(source, 'capacity', None)}, tagged {getattr(source, 'tagged_capacity', None)}, "
                                       f"slots {getattr(source, 'slot_geometry', None)})")
        try:
            chunks = _chunks(center.shape[0], source, streamed)
            sig = need = built = None
            if cacheable:
                # what a remembered list must have been built for: same system, same route, cutoffs at least today's
                need = need_tagged if tagged else need_plain
                sig = (str(dev), center.shape[1], tuple(float(b) for b in source.box), id(source.bonded), bool(tagged), bool(frame_route),
                       None if not tagged or source.tag[3] is None else id(source.tag[3]))
                built = _PAIR_LISTS.built_cutoffs(center, quat, sig)
                if built is None or any(n_ > b_ for n_, b_ in zip(need, built)):
                    built = tuple(x * (1.0 + PAIR_LIST_MARGIN) for x in need)  # (re)build with room for parameter updates
                if tagged:
                    source.tag = (source.tag[0], built[0], built[1], *source.tag[3:])
                else:
                    source.r_cutoff = built[0]
            if isinstance(source, CellListPairs):
                source.keep_lists = bool(cacheable)
            # streamed frames: the copies run ahead of the kernels on the copy stream, as many chunks as STREAM_AHEAD_BYTES
            # allows (the copy engine never idles; a chunk's kernels wait only for that chunk's own event)
            ahead, queued, in_flight = collections.deque(), 0, 0
            per_frame = (center[0].numel() + quat[0].numel()) * center.element_size() if streamed and center.shape[0] else 0

            def top_up():
                nonlocal queued, in_flight
                while queued < len(chunks):
                    nb = (chunks[queued].stop - chunks[queued].start) * per_frame
                    if ahead and in_flight + nb > STREAM_AHEAD_BYTES:
                        break
                    ahead.append((_fetch(center, quat, chunks[queued], dev), nb))
                    in_flight += nb
                    queued += 1

            if streamed:
                top_up()
            for k, sl in enumerate(chunks):
                if streamed:
                    (c_sl, q_sl, ready), nb = ahead.popleft()
                    in_flight -= nb
                    top_up()
                    torch.cuda.current_stream(dev).wait_event(ready)
                    c_sl.record_stream(torch.cuda.current_stream(dev))
                    q_sl.record_stream(torch.cuda.current_stream(dev))
                else:
                    c_sl, q_sl = center[sl], quat[sl]
                hit = _PAIR_LISTS.lookup(center, quat, sig, need, sl.start, sl.stop) if cacheable else None
                if hit is not None:
                    pairs, stride, count, split = hit
                else:
                    if tagged:
                        pairs, stride, count = source.chunk(sl, c_sl, q_sl, tagged=True, slots=frame_route)
                    else:
                        pairs, stride, count = source.chunk(sl, c_sl)
                    split = source.last_split if tagged else None
                    if cacheable and pairs is not None:
                        nbytes = pairs.numel() * 4 + (0 if count is None else count.numel() * 4)
                        source._cache_candidates.append((center, quat, sig, built, sl.start, sl.stop, (pairs, stride, count, split), nbytes))
                outs.append(
                    _launch(model, topo, c_sl, q_sl, params, pairs, stride, term_mask,
                            None if cot is None else cot[sl], want_terms, want_pos, want_par, per_frame_par, count,
                            flags | (_lib.FLAG_TAGGED_PAIRS if tagged else 0), 0.0, None, split,
                            observables=None if observables is None else (observables, observables.out[sl]), pseq=pseq)
                )
        except _lib.MythosB200Error as err:
            if not tagged or getattr(err, "status", None) != 3:  # MB_ECAPACITY: the frame-resident kernel does not apply
                raise
            tagged, source.tag = False, None
            source._cache_candidates.clear()
            source._pending.clear()
            source._extents.clear()
            source._slot_stats.clear()
            continue
        if isinstance(source, CellListPairs) and _DEFERRED:
            _DEFERRED[-1].sources.append(source)  # the caller reads the overflow flags after ITS launches (deferred_verification)
            break
        if not isinstance(source, CellListPairs) or source.verify():
            break
        if os.environ.get("MYTHOS_B200_DEBUG"):
            print("[mythos_b200] pair-list pass repeated:", source.slot_geometry, source.capacity, source.tagged_capacity, flush=True)
    return _merge(outs, want_terms, want_pos, want_par, per_frame_par)


_DEFERRED: list = []


class deferred_verification:  # noqa: N801 - used as a context manager
    """Inside this context the per-pass host read of the pair-list overflow flags is NOT done by the evaluation itself:
    the caller enqueues everything that consumes the result (weights, loss, collectives) first and then calls ``ok()``
    -- one host sync at the end instead of one in the middle that leaves the GPU idle while the rest is launched.
    ``ok()`` False means a list overflowed (capacities have been grown): the results are invalid, redo the pass."""

    def __init__(self):
        self.sources: list = []

    def __enter__(self):
        _DEFERRED.append(self)
        return self

    def __exit__(self, *exc):
        _DEFERRED.remove(self)
        if exc[0] is not None:
            for src in self.sources:
                src._cache_candidates.clear()
                src._pending.clear()
                src._extents.clear()
                src._slot_stats.clear()
                src._queued = None
        return False

    def enqueue(self) -> None:
        """Start the read-back of the pass's overflow flags (reductions + copy into pinned memory) without waiting for it;
        call it when the forward launches are enqueued, ``ok()`` after the step's own host sync."""
        for src in self.sources:
            src.enqueue_verification()

    def ok(self) -> bool:
        good = True
        for src in self.sources:
            good = src.verify() and good
        self.sources = []
        return good


_COPY_STREAMS: dict = {}  # per device
_PREFETCHED: dict = {}  # at most one entry: the first chunk of a pinned trajectory, copied ahead of its pass
_STREAM_LOCK = threading.Lock()  # guards the three module-level tables below against concurrent host threads


def prefetch_frames(center: torch.Tensor, quat: torch.Tensor) -> None:
    """Start the host -> device copy of the first chunk of pinned host frames NOW, before the caller's host-side work
    (the theta -> parameter-bank chain of a DiffTRe step), so the first kernels do not wait for PCIe.  A no-op for
    device tensors; the copy is picked up by the next streamed pass over the same buffers."""
    with _STREAM_LOCK:
        _PREFETCHED.clear()
    if center.is_cuda or not (center.is_pinned() and quat.is_pinned() and torch.cuda.is_available()) or center.dim() != 3:
        return
    dev = torch.device("cuda", torch.cuda.current_device())
    sl = _chunks(center.shape[0], CellListPairs, True)[0]  # as a pass with per-frame device lists (the DiffTRe route) cuts it
    key = (center.data_ptr(), quat.data_ptr(), center.shape[0], sl.start, sl.stop, dev)
    fetched = _fetch(center, quat, sl, dev)
    with _STREAM_LOCK:
        _PREFETCHED[key] = fetched


def _fetch(center: torch.Tensor, quat: torch.Tensor, sl: slice, dev):
    """Enqueue the host -> device copy of one chunk of frames on the device's copy stream; -> (center, quat, event)."""
    with _STREAM_LOCK:
        hit = None
        if _PREFETCHED:
            hit = _PREFETCHED.pop((center.data_ptr(), quat.data_ptr(), center.shape[0], sl.start, sl.stop, dev), None)
            _PREFETCHED.clear()
        stream = _COPY_STREAMS.get(dev)
        if stream is None:
            stream = _COPY_STREAMS[dev] = torch.cuda.Stream(device=dev)
    if hit is not None:
        return hit
    with torch.cuda.stream(stream):
        c = center[sl].to(dev, non_blocking=True)
        q = quat[sl].to(dev, non_blocking=True)
        ready = torch.cuda.Event()
        ready.record(stream)
    return c, q, ready


def _merge(outs, want_terms, want_pos, want_par, per_frame_par):
    if len(outs) == 1:
        return outs[0]
    terms = torch.cat([o[0] for o in outs]) if want_terms else None
    d_center = torch.cat([o[1] for o in outs]) if want_pos else None
    d_quat = torch.cat([o[2] for o in outs]) if want_pos else None
    d_params = None
    if want_par:
        d_params = torch.cat([o[3] for o in outs]) if per_frame_par else torch.stack([o[3] for o in outs]).sum(0)
    return terms, d_center, d_quat, d_params


class _EnergyTerms(torch.autograd.Function):
    """(center, quat, params) -> per-term energies (F,8); backward = second launch with the cotangent (remat)."""

    @staticmethod
    def forward(ctx, center, quat, params, model, topo, source, term_mask):
        terms, _, _, _ = _run(model, topo, center, quat, params, source, term_mask, None, True, False, False, False)
        ctx.save_for_backward(center, quat, params)
        ctx.static = (model, topo, source, term_mask)
        return terms

    @staticmethod
    def backward(ctx, g_terms):
        center, quat, params = ctx.saved_tensors
        model, topo, source, mask = ctx.static
        need_pos = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        need_par = ctx.needs_input_grad[2]
        _, d_center, d_quat, d_params = _run(
            model, topo, center, quat, params, source, mask, g_terms.contiguous(), False, need_pos, need_par, False
        )
        if d_params is not None:
            d_params = d_params.to(device=params.device, dtype=params.dtype)
        return (d_center if ctx.needs_input_grad[0] else None, d_quat if ctx.needs_input_grad[1] else None, d_params,
                None, None, None, None)


class _FrameEnergy(torch.autograd.Function):
    """(center, quat, params) -> weighted total energy per frame (F,), with dE/dparams rows produced in the SAME pass.

    This is the DiffTRe shape: positions are constants, the loss needs ``E_k`` and ``sum_k g_k dE_k/dparams``.
    When only ``params`` needs a gradient the forward launch also writes the Jacobian rows ``J (F, n_banks*P)``
    (one fused energy + parameter-gradient pass over the pair list), and the backward is the tiny product ``g @ J``.
    """

    @staticmethod
    def forward(ctx, center, quat, params, weights, model, topo, source, term_mask, observables=None):
        F = center.shape[0]
        cot = weights.to(device=params.device, dtype=center.dtype).reshape(1, -1).expand(F, -1).contiguous()
        jac_now = ctx.needs_input_grad[2] and not (ctx.needs_input_grad[0] or ctx.needs_input_grad[1])
        terms, _, _, J = _run(model, topo, center, quat, params, source, term_mask, cot, True, False, jac_now, True,
                              observables=observables)
        ctx.jac_now = jac_now
        if jac_now:
            ctx.save_for_backward(J)
        else:
            ctx.save_for_backward(center, quat, params, cot)
        ctx.static = (model, topo, source, term_mask, params.device, params.dtype)
        return (terms * cot).sum(dim=1)

    @staticmethod
    def backward(ctx, g):
        model, topo, source, mask, pdev, pdtype = ctx.static
        if ctx.jac_now:
            (J,) = ctx.saved_tensors
            return None, None, (g.to(J.dtype) @ J).to(device=pdev, dtype=pdtype), None, None, None, None, None, None
        center, quat, params, cot = ctx.saved_tensors
        need_pos = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        _, d_center, d_quat, d_params = _run(
            model, topo, center, quat, params, source, mask, (cot * g.reshape(-1, 1)).contiguous(), False, need_pos,
            ctx.needs_input_grad[2], False,
        )
        if d_params is not None:
            d_params = d_params.to(device=pdev, dtype=pdtype)
        return (d_center if ctx.needs_input_grad[0] else None, d_quat if ctx.needs_input_grad[1] else None, d_params,
                None, None, None, None, None, None)


@dc.dataclass
class PseqInputs:
    """Device-side inputs of the probabilistic-sequence weights (``mb_pseq``); ``grads=True`` adds zeroed gradient buffers."""

    pmarg: torch.Tensor  # (N,4)
    bp_of: torch.Tensor  # (N) int32
    within: torch.Tensor  # (N) int32
    same_w_stack: torch.Tensor  # (n_bp,2)
    same_w_hb: torch.Tensor  # (n_bp,2)
    terms: int
    d_pmarg: torch.Tensor | None = None
    d_same_w_stack: torch.Tensor | None = None
    d_same_w_hb: torch.Tensor | None = None

    def with_grads(self) -> "PseqInputs":
        return dc.replace(self, d_pmarg=torch.zeros_like(self.pmarg), d_same_w_stack=torch.zeros_like(self.same_w_stack),
                          d_same_w_hb=torch.zeros_like(self.same_w_hb))

    def struct(self, dtype) -> _lib.Pseq:
        for t in (self.pmarg, self.same_w_stack, self.same_w_hb):
            if t.dtype != dtype or not t.is_contiguous():
                raise _lib.MythosB200Error("pseq inputs must be contiguous and of the frames' dtype")
        s = _lib.Pseq()
        s.pmarg, s.bp_of, s.within = self.pmarg.data_ptr(), self.bp_of.data_ptr(), self.within.data_ptr()
        s.same_w_stack = self.same_w_stack.data_ptr() if self.same_w_stack.numel() else None
        s.same_w_hb = self.same_w_hb.data_ptr() if self.same_w_hb.numel() else None
        s.d_pmarg = _lib.ptr(self.d_pmarg)
        s.d_same_w_stack = self.d_same_w_stack.data_ptr() if self.d_same_w_stack is not None and self.d_same_w_stack.numel() else None
        s.d_same_w_hb = self.d_same_w_hb.data_ptr() if self.d_same_w_hb is not None and self.d_same_w_hb.numel() else None
        s.terms = self.terms
        return s


class _PseqTerms(torch.autograd.Function):
    """Per-term energies with probabilistic sequence weights: (center, quat, params, pmarg, same_w_stack, same_w_hb) ->
    (F,8); the backward is a second launch that also returns the gradients with respect to the three weight inputs."""

    @staticmethod
    def forward(ctx, center, quat, params, pmarg, same_s, same_h, model, topo, source, term_mask, bp_of, within, terms):
        dev, dtype = center.device, center.dtype
        ps = PseqInputs(pmarg.to(dev, dtype).contiguous(), bp_of, within, same_s.to(dev, dtype).contiguous(),
                        same_h.to(dev, dtype).contiguous(), terms)
        out, _, _, _ = _run(model, topo, center, quat, params, source, term_mask, None, True, False, False, False, pseq=ps)
        ctx.save_for_backward(center, quat, params, pmarg, same_s, same_h)
        ctx.static = (model, topo, source, term_mask, bp_of, within, terms)
        return out

    @staticmethod
    def backward(ctx, g_terms):
        center, quat, params, pmarg, same_s, same_h = ctx.saved_tensors
        model, topo, source, mask, bp_of, within, terms = ctx.static
        dev, dtype = center.device, center.dtype
        ps = PseqInputs(pmarg.to(dev, dtype).contiguous(), bp_of, within, same_s.to(dev, dtype).contiguous(),
                        same_h.to(dev, dtype).contiguous(), terms).with_grads()
        need_pos = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        _, d_center, d_quat, d_params = _run(model, topo, center, quat, params, source, mask, g_terms.contiguous(), False, need_pos,
                                             True, False, pseq=ps)
        back = lambda g, like: g.to(device=like.device, dtype=like.dtype)  # noqa: E731
        return (d_center if ctx.needs_input_grad[0] else None, d_quat if ctx.needs_input_grad[1] else None,
                back(d_params, params), back(ps.d_pmarg, pmarg), back(ps.d_same_w_stack, same_s), back(ps.d_same_w_hb, same_h),
                None, None, None, None, None, None, None)


def pseq_energy_terms(model, topo, center, quat, params, pairs, term_mask, pmarg, same_w_stack, same_w_hb, bp_of, within, terms):
    """``energy_terms`` with probabilistic sequence weights (generic pair kernel)."""
    if center.dim() != 3 or quat.dim() != 3:
        raise _lib.MythosB200Error("center must be (F,N,3) and quat (F,N,4)")
    return _PseqTerms.apply(center, quat, params, pmarg, same_w_stack, same_w_hb, model, topo, _source_of(pairs), term_mask,
                            bp_of, within, terms)


def _source_of(pairs) -> StaticPairs | CellListPairs:
    return pairs if isinstance(pairs, (StaticPairs, CellListPairs)) else StaticPairs(pairs)


def energy_terms(model, topo, center, quat, params, pairs, term_mask: int = _lib.ALL_TERMS) -> torch.Tensor:
    """Per-term energies ``(F, 8)`` for ``center (F,N,3)``, ``quat (F,N,4)``; differentiable in center, quat, params.

    ``pairs``: a (2,U) / (F,2,U) int32 device tensor, ``StaticPairs`` or ``CellListPairs``."""
    if center.dim() != 3 or quat.dim() != 3:
        raise _lib.MythosB200Error("center must be (F,N,3) and quat (F,N,4)")
    return _EnergyTerms.apply(center, quat, params, model, topo, _source_of(pairs), term_mask)


def frame_energies(model, topo, center, quat, params, pairs, weights: torch.Tensor, term_mask: int = _lib.ALL_TERMS,
                   observables=None) -> torch.Tensor:
    """``sum_t weights[t] * E_t`` per frame, ``(F,)``; see ``_FrameEnergy`` for the fused parameter-gradient pass.
    ``observables``: an ``ObservableRequest`` (mythos_b200.observables.base) filled in the same pass."""
    if center.dim() != 3 or quat.dim() != 3:
        raise _lib.MythosB200Error("center must be (F,N,3) and quat (F,N,4)")
    return _FrameEnergy.apply(center, quat, params, weights, model, topo, _source_of(pairs), term_mask, observables)


def energy_and_gradients(model, topo, center, quat, params, pairs, cot=None, term_mask: int = _lib.ALL_TERMS,
                         want_pos_grad: bool = True, want_param_grad: bool = False, per_frame_param_grad: bool = False,
                         flags: int = 0):
    """One fused pass returning ``(terms, d_center, d_quat, d_params)`` without autograd bookkeeping.

    This is the call the MD loop and the DiffTRe pass use: energies, forces and the parameter gradient of
    ``sum_t cot[f,t] * E_t(frame f)`` come out of a single pass over the pair list."""
    return _run(model, topo, center, quat, params.to(device=center.device, dtype=center.dtype), _source_of(pairs), term_mask,
                cot, True, want_pos_grad, want_param_grad, per_frame_param_grad, flags)
