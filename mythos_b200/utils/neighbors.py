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
    fn = getattr(_lib.lib(), f"mythos_b200_nl_build_{_lib.suffix(center.dtype)}")
    with torch.cuda.device(dev):
        _lib.check(fn(_lib.current_stream(dev), C.byref(a)), "mythos_b200_nl_build")
    return pairs, count, overflow, workspace


@dc.dataclass
class NeighborList:
    """State of one neighbour list (fields named as in jax_md.partition.NeighborList)."""

    idx: torch.Tensor  # (2, capacity) int32, padded with N
    reference_position: torch.Tensor  # (N,3) centres at the last rebuild
    did_buffer_overflow: torch.Tensor  # (1,) int32 flag on the device (bit 0: capacity, bit 1: > 4 bonds on one nucleotide)
    count: torch.Tensor  # (1,) pairs found at the last rebuild
    fns: "NeighborListFns"
    workspace: torch.Tensor | None = None

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

    def _centers(self, position) -> torch.Tensor:
        c = position.center if hasattr(position, "center") else position
        return c

    def allocate(self, position, extra_capacity: int = 0) -> NeighborList:
        """Size the list from the current configuration (host sync, like jax_md's allocate) and build it."""
        c = self._centers(position)
        probe, count, overflow, ws = build_pairs(c.unsqueeze(0), self.bonded_neighbors, self.box, self.r_cutoff, self.dr_threshold, 1)
        n_found = int(count.item())
        capacity = (max(int(n_found * self.capacity_multiplier) + extra_capacity, 1) + 3) // 4 * 4  # multiple of 4: 128-bit pair stores
        pairs, count, overflow, ws = build_pairs(
            c.unsqueeze(0), self.bonded_neighbors, self.box, self.r_cutoff, self.dr_threshold, capacity, ws
        )
        return NeighborList(idx=pairs[0], reference_position=c.detach().clone(), did_buffer_overflow=overflow, count=count, fns=self, workspace=ws)

    def update(self, position, nbrs: NeighborList, force_rebuild: bool = False) -> NeighborList:
        """Rebuild if any nucleotide moved more than dr_threshold/2 since the last build (checked on the device;
        the rebuild itself is enqueued unconditionally when ``force_rebuild`` or dr_threshold == 0)."""
        c = self._centers(position)
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
