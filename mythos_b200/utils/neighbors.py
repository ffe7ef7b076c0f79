"""Neighbour lists of unbonded pairs, built on the device by the cell-list kernels.

Interface of ``mythos/utils/neighbors.py:12-59`` (``get_neighbor_list_fn(bonded_neighbors, n_nucleotides,
displacement_fn, box_size, r_cutoff=10.0, dr_threshold=0.2)`` returning an object with ``allocate`` / ``update``)
and of the ``jax_md.partition.NeighborList`` objects it hands out (``idx`` of shape ``(2, capacity)`` padded with N,
``reference_position``, ``did_buffer_overflow``, ``update(position)``).  The reference's list is an O(N^2)
distance matrix per rebuild (``disable_cell_list=True``); here it is integer cell binning + counting sort +
count/scan/fill (``mythos_b200/csrc/neighbors.cu``) with the same accept test, so the pair set is identical.
"""

from __future__ import annotations

import ctypes as C
import dataclasses as dc

import numpy as np
import torch

from mythos_b200 import _lib, space

ERR_NEIGHBORS_INVALID_BONDED_NEIGHBORS = "Indices of bonded neighbors must be bewteen 0 and n_nucleotides-1"


def build_pairs(
    center: torch.Tensor,
    bonded: torch.Tensor,
    box: tuple[float, float, float],
    r_cutoff: float,
    dr_threshold: float,
    capacity: int,
    workspace: torch.Tensor | None = None,
    rows: bool = False,
    max_row: torch.Tensor | None = None,
    tag_bits: int = 0,
    out: tuple | None = None,
    append_count: torch.Tensor | None = None,
    warp_slots: tuple[int, int, int] | None = None,
    reference: torch.Tensor | None = None,
    move_threshold: float = 0.0,
    rebuilds: torch.Tensor | None = None,
    reuse_exclusions: bool = False,
    packed_slots: bool = False,
) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """Raw batched build.  center (F,N,3) on the device -> (pairs (F,2,capacity) int32, count (F,), overflow (1,), workspace).

    ``rows=True`` is the one-pass layout (``MB_NL_ROWS``): fixed-width rows per nucleotide, unused slots = N; pass a
    ``max_row`` (F,) int32 tensor to receive the longest row of each frame."""
    _lib.require_cuda(center, "center")
    if center.dim() != 3:
        raise _lib.MythosB200Error("center must be (F,N,3)")
    F, N = center.shape[0], center.shape[1]
    dev = center.device
    center = center.contiguous()
    need = int(_lib.lib().mythos_b200_nl_workspace_bytes(N, F))
    if workspace is None or workspace.numel() < need or workspace.device != dev:
        workspace = torch.empty(need, dtype=torch.uint8, device=dev)
    if out is not None:  # caller-owned (pairs, count, overflow): a second, tagged build appending to the same list
        pairs, count, overflow = out
    else:
        pairs = torch.empty((F, 2, capacity), dtype=torch.int32, device=dev)
        count = torch.empty((F,), dtype=torch.int32, device=dev)
        overflow = torch.zeros((1,), dtype=torch.int32, device=dev)
    bonded = bonded.to(device=dev, dtype=torch.int32).contiguous().reshape(-1, 2)
    a = _lib.NlArgs()
    a.n, a.n_frames = N, F
    a.center = center.data_ptr()
    a.bonded = bonded.data_ptr() if bonded.numel() else None
    a.n_bonded = bonded.shape[0]
    for d in range(3):
        a.box[d] = float(box[d])
    a.r_cutoff, a.dr_threshold = float(r_cutoff), float(dr_threshold)
    a.pairs, a.capacity = pairs.data_ptr(), capacity
    a.count, a.overflow = count.data_ptr(), overflow.data_ptr()
    a.workspace, a.workspace_bytes = workspace.data_ptr(), workspace.numel()
    a.flags = _lib.NL_ROWS if rows else 0
    a.max_row = max_row.data_ptr() if (rows and max_row is not None) else None
    if warp_slots is not None:  # one-pass layout: (lane_slots, slot_base, slot_width); max_row is (F,2)
        a.flags |= _lib.NL_WARP_SLOTS
        a.lane_slots, a.slot_base, a.slot_width = int(warp_slots[0]), int(warp_slots[1]), int(warp_slots[2])
        a.max_row = max_row.data_ptr() if max_row is not None else None
    if tag_bits or append_count is not None:  # support tags (internal contract with this library's energy kernels)
        a.flags |= _lib.NL_TAG_SUPPORTS
        a.tag_bits = int(tag_bits)
        a.append_count = append_count.data_ptr() if append_count is not None else None
    if reference is not None:  # conditional rebuild on the device: (F,N,3) centres at the last build, in/out
        if reference.shape != center.shape or reference.dtype != center.dtype or not reference.is_contiguous() or reference.device != dev:
            raise _lib.MythosB200Error("reference must be a contiguous (F,N,3) tensor like center")
        a.reference, a.move_threshold = reference.data_ptr(), float(move_threshold)
        a.rebuilds = rebuilds.data_ptr() if rebuilds is not None else None
    if reuse_exclusions:
        a.flags |= _lib.NL_REUSE_EXCLUSIONS
    if packed_slots:  # warp slots packed back to back: a compact list, count[f] valid entries at its head (frame-resident route only)
        a.flags |= _lib.NL_PACKED_SLOTS
    fn = getattr(_lib.lib(), f"mythos_b200_nl_build_{_lib.suffix(center.dtype)}")
    with torch.cuda.device(dev):
        _lib.check(fn(_lib.current_stream(dev), C.byref(a)), "mythos_b200_nl_build")
    return pairs, count, overflow, workspace


DEVICE_SIDE_UPDATE = True  # False: always the compact list + host-checked update (the reference's control flow)


@dc.dataclass
class NeighborList:
    """State of one neighbour list (fields named as in jax_md.partition.NeighborList)."""

    idx: torch.Tensor  # (2, capacity) int32, padded with N
    reference_position: torch.Tensor  # (N,3) centres at the last rebuild
    did_buffer_overflow: torch.Tensor  # (1,) int32 flag on the device (bit 0: capacity, bit 1: > 4 bonds on one nucleotide)
    count: torch.Tensor  # (1,) pairs found at the last rebuild
    fns: "NeighborListFns"
    workspace: torch.Tensor | None = None
    # device-side update (free space, small systems): the list lives in the one-pass warp-slot layout (a padded OrderedSparse
    # list like any other) and update() is ONE conditional launch that rebuilds it in place -- no host round trip per step
    slots: tuple[int, int] | None = None  # (lane_slots, slot_width)
    rebuilds: torch.Tensor | None = None  # (1,) int32 on the device: rebuilds since allocate()
    max_row: torch.Tensor | None = None  # (1,2) int32: longest lane row / warp total of the last rebuild

    def update(self, position: torch.Tensor, force_rebuild: bool = False) -> "NeighborList":
        return self.fns.update(position, self, force_rebuild=force_rebuild)


@dc.dataclass
class NeighborListFns:
    bonded_neighbors: torch.Tensor
    n_nucleotides: int
    box: tuple[float, float, float]
    r_cutoff: float
    dr_threshold: float
    capacity_multiplier: float = 1.25

    def _bonded_on(self, device) -> torch.Tensor:
        """The bonded list on the device (cached: update() runs every MD step, also while a CUDA graph is captured)."""
        cache = self.__dict__.setdefault("_bonded_dev", {})
        b = cache.get(str(device))
        if b is None:
            b = cache[str(device)] = self.bonded_neighbors.to(device=device, dtype=torch.int32).contiguous().reshape(-1, 2)
        return b

    def _centers(self, position) -> torch.Tensor:
        c = position.center if hasattr(position, "center") else position
        return c

    def allocate(self, position, extra_capacity: int = 0) -> NeighborList:
        """Size the list from the current configuration (host sync, like jax_md's allocate) and build it."""
        c = self._centers(position)
        nl = self._allocate_on_device(c, extra_capacity)
        if nl is not None:
            return nl
        probe, count, overflow, ws = build_pairs(c.unsqueeze(0), self.bonded_neighbors, self.box, self.r_cutoff, self.dr_threshold, 1)
        n_found = int(count.item())
        capacity = (max(int(n_found * self.capacity_multiplier) + extra_capacity, 1) + 3) // 4 * 4  # multiple of 4: 128-bit pair stores
        pairs, count, overflow, ws = build_pairs(
            c.unsqueeze(0), self.bonded_neighbors, self.box, self.r_cutoff, self.dr_threshold, capacity, ws
        )
        return NeighborList(idx=pairs[0], reference_position=c.detach().clone(), did_buffer_overflow=overflow, count=count, fns=self, workspace=ws)

    def _allocate_on_device(self, c: torch.Tensor, extra_capacity: int) -> NeighborList | None:
        """The list in the warp-slot layout with a device-side conditional update, where the frame-resident build applies
        (free space, cell table and records of the system fit in shared memory); None otherwise."""
        if not DEVICE_SIDE_UPDATE or any(self.box) or not c.is_cuda or c.dim() != 2:
            return None
        with torch.cuda.device(c.device):
            if not _lib.lib().mythos_b200_nl_conditional_supported(c.shape[0], 256, c.element_size()):
                return None
        n = c.shape[0]
        wpf = (n + 31) // 32
        cc = c.detach().contiguous().unsqueeze(0)
        mr = torch.zeros((1, 2), dtype=torch.int32, device=c.device)
        _, _, _, ws = build_pairs(cc, self._bonded_on(c.device), self.box, self.r_cutoff, self.dr_threshold, wpf * 32 * 256, None,
                                  max_row=mr, warp_slots=(256, 0, 32 * 256))  # probe with the widest slots: sizes only
        lane_max, warp_max = (int(v) for v in mr[0].tolist())
        lane_slots = min(int(lane_max * self.capacity_multiplier) + 4, 256)
        slot_width = (int(warp_max * self.capacity_multiplier) + -(-extra_capacity // wpf) + 16 + 3) // 4 * 4
        capacity = wpf * slot_width
        pairs = torch.empty((1, 2, capacity), dtype=torch.int32, device=c.device)
        count = torch.zeros((1,), dtype=torch.int32, device=c.device)
        overflow = torch.zeros((1,), dtype=torch.int32, device=c.device)
        build_pairs(cc, self._bonded_on(c.device), self.box, self.r_cutoff, self.dr_threshold, capacity, ws, max_row=mr,
                    out=(pairs, count, overflow), warp_slots=(lane_slots, 0, slot_width))
        return NeighborList(idx=pairs[0], reference_position=cc.clone()[0], did_buffer_overflow=overflow, count=count, fns=self,
                            workspace=ws, slots=(lane_slots, slot_width), rebuilds=torch.zeros((1,), dtype=torch.int32, device=c.device),
                            max_row=mr)

    def update(self, position, nbrs: NeighborList, force_rebuild: bool = False) -> NeighborList:
        """Rebuild if any nucleotide moved more than dr_threshold/2 since the last build (checked on the device;
        the rebuild itself is enqueued unconditionally when ``force_rebuild`` or dr_threshold == 0)."""
        c = self._centers(position)
        if nbrs.slots is not None:
            # one launch: displacement test, reference update and rebuild all happen on the device, in place (the object
            # handed back is the same one; bit 0 / 2 of did_buffer_overflow report slots that became too small)
            cc = c.detach().contiguous().unsqueeze(0)
            capacity = nbrs.idx.shape[-1]
            build_pairs(cc, self._bonded_on(c.device), self.box, self.r_cutoff, self.dr_threshold, capacity, nbrs.workspace,
                        max_row=nbrs.max_row, out=(nbrs.idx.unsqueeze(0), nbrs.count, nbrs.did_buffer_overflow),
                        warp_slots=(nbrs.slots[0], 0, nbrs.slots[1]),
                        reference=None if force_rebuild else nbrs.reference_position.unsqueeze(0),
                        move_threshold=0.5 * self.dr_threshold, rebuilds=nbrs.rebuilds, reuse_exclusions=True)
            if force_rebuild:
                nbrs.reference_position.copy_(cc[0])
            return nbrs
        if not force_rebuild and self.dr_threshold > 0:
            d = space.Displacement(self.box if any(self.box) else None)(c, nbrs.reference_position)
            moved = (d * d).sum(-1).max() > (0.5 * self.dr_threshold) ** 2
            if not bool(moved.item()):
                return nbrs
        capacity = nbrs.idx.shape[-1]
        pairs, count, overflow, ws = build_pairs(
            c.unsqueeze(0), self.bonded_neighbors, self.box, self.r_cutoff, self.dr_threshold, capacity, nbrs.workspace
        )
        overflow |= nbrs.did_buffer_overflow
        return NeighborList(idx=pairs[0], reference_position=c.detach().clone(), did_buffer_overflow=overflow, count=count, fns=self, workspace=ws)


def get_neighbor_list_fn(
    bonded_neighbors,
    n_nucleotides: int,
    displacement_fn,
    box_size=None,
    r_cutoff: float = 10.0,
    dr_threshold: float = 0.2,
    capacity_multiplier: float = 1.25,
) -> NeighborListFns:
    """Neighbour-list factory for unbonded pairs (bonded pairs are excluded from the list)."""
    b = np.asarray(bonded_neighbors.cpu() if isinstance(bonded_neighbors, torch.Tensor) else bonded_neighbors)
    if not ((b >= 0) & (b < n_nucleotides)).all():
        raise ValueError(ERR_NEIGHBORS_INVALID_BONDED_NEIGHBORS)
    # the box that matters is the one the displacement applies (free space -> no wrap, whatever box_size says;
    # jax_md only uses box_size to dimension its cell grid, which this build derives from the data)
    box = space.box_of(displacement_fn)
    return NeighborListFns(
        bonded_neighbors=torch.as_tensor(b, dtype=torch.int32).reshape(-1, 2),
        n_nucleotides=int(n_nucleotides),
        box=box,
        r_cutoff=float(r_cutoff),
        dr_threshold=float(dr_threshold),
        capacity_multiplier=capacity_multiplier,
    )
