"""Sequence constraints and probabilistic sequences (interface of ``mythos/input/sequence_constraints.py:77-230``).

A probabilistic sequence is ``(unpaired_pseq (n_unpaired,4), bp_pseq (n_bp,4))``: a nucleotide distribution per unpaired
nucleotide and a distribution over the four base-pair types ``AT, TA, GC, CG`` per declared base pair."""

from __future__ import annotations

import dataclasses as dc

import numpy as np
import torch

DNA_ALPHA = "ACGT"
BP_TYPES = ["AT", "TA", "GC", "CG"]  # mythos/utils/constants.py:13
BP_IDXS = np.array([[DNA_ALPHA.index(a), DNA_ALPHA.index(b)] for a, b in BP_TYPES])  # constants.py:18
BP_IDX_MAP = {(int(a), int(b)): k for k, (a, b) in enumerate(BP_IDXS)}

ERR_SEQ_CONSTRAINTS_INVALID_NUMBER_NUCLEOTIDES = "Invalid number of nucleotides"
ERR_INVALID_BP_SHAPE = "Invalid shape for base pairs"
ERR_SEQ_CONSTRAINTS_MISMATCH_NUM_TYPES = (
    "Number of nucleotides should equal the number of unpaired base pairs plus the number of coupled base pairs"
)
ERR_SEQ_CONSTRAINTS_INVALID_COVER = "Unpaired and coupled nucleotides do not cover all nucleotides"
ERR_BP_ARR_CONTAINS_DUPLICATES = "Array specifying base paired indices cannot contain duplicates"
ERR_INVALID_BP_INDICES = "Base paired indices must be between 0 and n_nucleotides-1"
ERR_DSEQ_TO_PSEQ_INVALID_BP = "Invalid base pair encountered when converting discrete sequence to probabilistic sequence"


@dc.dataclass(frozen=True)
class SequenceConstraints:
    """Which nucleotides are unpaired and which form base pairs, with the index maps the weights need."""

    n_nucleotides: int
    n_unpaired: int
    n_bp: int
    is_unpaired: np.ndarray  # (N) 0/1
    unpaired: np.ndarray  # (n_unpaired)
    bps: np.ndarray  # (n_bp,2)
    idx_to_unpaired_idx: np.ndarray  # (N) or -1
    idx_to_bp_idx: np.ndarray  # (N,2): (base-pair index, position inside it) or (-1,-1)

    def __post_init__(self) -> None:
        if self.n_nucleotides < 1:
            raise ValueError(ERR_SEQ_CONSTRAINTS_INVALID_NUMBER_NUCLEOTIDES)
        if np.asarray(self.bps).shape != (self.n_bp, 2):
            raise ValueError(ERR_INVALID_BP_SHAPE)
        if self.n_unpaired + 2 * self.n_bp != self.n_nucleotides:
            raise ValueError(ERR_SEQ_CONSTRAINTS_MISMATCH_NUM_TYPES)
        cover = set(np.concatenate([np.asarray(self.unpaired).reshape(-1), np.asarray(self.bps).reshape(-1)]).tolist())
        if cover != set(range(self.n_nucleotides)):
            raise ValueError(ERR_SEQ_CONSTRAINTS_INVALID_COVER)


def from_bps(n_nucleotides: int, bps) -> SequenceConstraints:
    """Constraints from a list of base pairs; every other nucleotide is unpaired (``sequence_constraints.py:130-177``)."""
    bps = np.asarray(bps, dtype=np.int64)
    if bps.ndim != 2 or bps.shape[1] != 2 or 2 * bps.shape[0] > n_nucleotides:
        raise ValueError(ERR_INVALID_BP_SHAPE)
    paired = bps.reshape(-1)
    if len(np.unique(paired)) < len(paired):
        raise ValueError(ERR_BP_ARR_CONTAINS_DUPLICATES)
    if not np.all((paired >= 0) & (paired < n_nucleotides)):
        raise ValueError(ERR_INVALID_BP_INDICES)
    unpaired = np.setdiff1d(np.arange(n_nucleotides), paired)
    idx_to_unpaired_idx = np.full((n_nucleotides,), -1, dtype=np.int32)
    idx_to_unpaired_idx[unpaired] = np.arange(len(unpaired), dtype=np.int32)
    idx_to_bp_idx = np.full((n_nucleotides, 2), -1, dtype=np.int32)
    for k, (a, b) in enumerate(bps):
        idx_to_bp_idx[a] = [k, 0]
        idx_to_bp_idx[b] = [k, 1]
    is_unpaired = np.zeros(n_nucleotides, dtype=np.int32)
    is_unpaired[unpaired] = 1
    return SequenceConstraints(n_nucleotides=n_nucleotides, n_unpaired=len(unpaired), n_bp=bps.shape[0], is_unpaired=is_unpaired,
                               unpaired=unpaired, bps=bps, idx_to_unpaired_idx=idx_to_unpaired_idx, idx_to_bp_idx=idx_to_bp_idx)


def dseq_to_pseq(dseq, sc: SequenceConstraints):
    """One-hot probabilistic sequence of a discrete one (``sequence_constraints.py:180-213``)."""
    dseq = np.asarray(dseq.cpu() if isinstance(dseq, torch.Tensor) else dseq)
    up = np.zeros((sc.n_unpaired, 4))
    for k, idx in enumerate(sc.unpaired):
        up[k, int(dseq[idx])] = 1.0
    bp = np.zeros((max(sc.n_bp, 1), 4))  # (the reference keeps a dummy row when there are no base pairs)
    for k, (a, b) in enumerate(sc.bps):
        key = (int(dseq[a]), int(dseq[b]))
        if key not in BP_IDX_MAP:
            raise ValueError(ERR_DSEQ_TO_PSEQ_INVALID_BP)
        bp[k, BP_IDX_MAP[key]] = 1.0
    return torch.tensor(up), torch.tensor(bp)
