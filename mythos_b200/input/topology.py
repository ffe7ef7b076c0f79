"""oxDNA topology files -> the arrays the energy path consumes.

Same contract as ``mythos/input/topology.py:86-327``: classic (``N n_strands`` header, one nucleotide per line) and
new (``N n_strands 5->3`` header, one strand per line) formats; nucleotides are kept in the internal 3'->5' order (the
new format's strands are reversed on read); bonded pairs are consecutive indices per strand plus ``(first, last)``
for circular strands; ``is_end`` flags strand ends; ``nt_type`` is 1 = DNA, 2 = RNA, 0 = unspecified.
``unbonded_neighbors`` -- the reference's all-pairs-minus-bonded ``(U,2)`` list -- is built lazily with numpy and
only for small systems; large systems use the cell-list neighbour build instead.
"""

from __future__ import annotations

import dataclasses as dc
from enum import IntEnum
from functools import cached_property
from pathlib import Path

import numpy as np

NUCLEOTIDES_IDX = {"A": 0, "C": 1, "G": 2, "T": 3, "U": 3}
ALL_PAIRS_LIMIT = 20000  # nucleotides; beyond this an all-pairs list is > 1.6 GB of int32 pairs
ALL_PAIRS_EXPLICIT = 512  # above this Topology.unbonded_neighbors is the AllPairs sentinel


class NucleotideType(IntEnum):
    UNSPECIFIED = 0
    DNA = 1
    RNA = 2


class AllPairs:
    """Sentinel for "every i<j that is not bonded" (what ``topology.unbonded_neighbors`` means in the reference,
    ``topology.py:186-190``) without materialising the O(N^2) list.  Energy functions holding it evaluate the
    unbonded terms over per-frame device cell lists at the interaction range of their parameters -- the same
    energies, since every unbonded term has compact support."""

    def __init__(self, n: int, in_kernel: bool = False):
        self.n = int(n)
        # True: where the frame-resident kernel applies, let it find the pairs itself (shared-memory cell list) instead
        # of streaming device lists through it
        self.in_kernel = bool(in_kernel)

    @property
    def T(self) -> "AllPairs":  # noqa: N802 - BaseEnergyFunction stores topology.unbonded_neighbors.T
        return self

    def __repr__(self) -> str:
        return f"AllPairs(n={self.n})"


def bonded_pairs(strand_counts, is_circular=None) -> np.ndarray:
    out, start = [], 0
    for k, n in enumerate(strand_counts):
        n = int(n)
        out += [(a, a + 1) for a in range(start, start + n - 1)]
        if is_circular is not None and is_circular[k]:
            out.append((start, start + n - 1))  # the ordering is the reference's (topology.py:178-180)
        start += n
    return np.array(out, dtype=np.int32).reshape(-1, 2)


def unbonded_pairs(n: int, bonded: np.ndarray) -> np.ndarray:
    """All i<j that are not bonded, (U,2) int32 (``topology.py:186-190``; row order is not significant)."""
    if n > ALL_PAIRS_LIMIT:
        raise ValueError(
            f"an all-pairs list for {n} nucleotides is not materialised; build a neighbour list "
            "(mythos_b200.utils.neighbors) and pass it with with_props(unbonded_neighbors=...)"
        )
    i, j = np.triu_indices(n, k=1)
    keep = np.ones(i.shape[0], dtype=bool)
    if bonded.size:
        lo, hi = np.minimum(bonded[:, 0], bonded[:, 1]).astype(np.int64), np.maximum(bonded[:, 0], bonded[:, 1]).astype(np.int64)
        keep &= ~np.isin(i.astype(np.int64) * n + j, lo * n + hi)
    return np.stack([i[keep], j[keep]], axis=1).astype(np.int32)


@dc.dataclass(frozen=True)
class Topology:
    n_nucleotides: int
    strand_counts: np.ndarray
    bonded_neighbors: np.ndarray
    seq: np.ndarray
    is_end: np.ndarray
    nt_type: np.ndarray
    is_circular: np.ndarray | None = None
    unbonded_override: np.ndarray | None = None

    def __post_init__(self) -> None:
        if self.n_nucleotides < 1:
            raise ValueError("Invalid number of nucleotides")
        if len(self.strand_counts) == 0 or sum(self.strand_counts) == 0:
            raise ValueError("Invalid strand counts")
        if self.n_nucleotides != sum(self.strand_counts):
            raise ValueError("Strand counts do not match number of nucleotides")
        if self.bonded_neighbors.ndim != 2 or self.bonded_neighbors.shape[1] != 2:
            raise ValueError("Invalid bonded neighbors shape")
        if self.seq.shape != (self.n_nucleotides,) or len(set(np.asarray(self.seq).tolist()) - {0, 1, 2, 3}) > 0:
            raise ValueError("Invalid discrete sequence")

    @cached_property
    def unbonded_neighbors(self):
        """(U,2) all-pairs-minus-bonded list for small systems, the ``AllPairs`` sentinel beyond ``ALL_PAIRS_EXPLICIT``."""
        if self.unbonded_override is not None:
            return self.unbonded_override
        if self.n_nucleotides > ALL_PAIRS_EXPLICIT:
            return AllPairs(self.n_nucleotides)
        return unbonded_pairs(self.n_nucleotides, self.bonded_neighbors)

    @cached_property
    def unbonded_neighbors_t(self):
        """The (2,U) orientation the energy functions store -- one shared object, so that all terms built from this
        topology are recognised as using the same list and fuse into one launch."""
        ub = self.unbonded_neighbors
        return ub if isinstance(ub, AllPairs) else np.ascontiguousarray(ub.T)


def from_strands(sequences: list[str], nt_types: list[int] | None = None, circular: list[bool] | None = None) -> Topology:
    """Topology from per-strand sequences given in the internal 3'->5' order."""
    counts = np.array([len(s) for s in sequences], dtype=np.int32)
    circ = np.array(circular if circular is not None else [False] * len(sequences), dtype=bool)
    seq = np.array([NUCLEOTIDES_IDX[c] for s in sequences for c in s], dtype=np.int32)
    is_end, nt = [], []
    for k, s in enumerate(sequences):
        e = [0] * len(s)
        if not circ[k] and len(s):
            e[0] = e[-1] = 1
        is_end += e
        if nt_types is not None:
            t = nt_types[k]
        else:
            t = NucleotideType.DNA if "T" in s else NucleotideType.RNA if "U" in s else NucleotideType.UNSPECIFIED
        nt += [int(t)] * len(s)
    return Topology(
        n_nucleotides=int(counts.sum()), strand_counts=counts, bonded_neighbors=bonded_pairs(counts, circ), seq=seq,
        is_end=np.array(is_end, dtype=np.int32), nt_type=np.array(nt, dtype=np.int32), is_circular=circ,
    )


def from_oxdna_file(path, *, return_format: bool = False):
    path = Path(path)
    if not path.exists():
        raise FileNotFoundError("Topology file not found")
    lines = [ln for ln in path.read_text().splitlines() if ln.strip()]
    head = lines[0].split()
    if len(head) == 2:
        fmt = "classic"
        rows = [ln.split() for ln in lines[1:]]
        seqs, circ = [], []
        for sid in sorted({int(r[0]) for r in rows}):
            sel = [r for r in rows if int(r[0]) == sid]
            seqs.append("".join(r[1] for r in sel))
            circ.append(int(sel[-1][3]) != -1)
        top = from_strands(seqs, circular=circ)
    elif len(head) == 3:
        fmt = "new"
        seqs, circ, types = [], [], []
        for ln in lines[1:]:
            seqs.append(ln.split()[0][::-1])  # file is 5'->3'; internal order is 3'->5'
            circ.append("circular=true" in ln)
            types.append(1 if "type=DNA" in ln else 2 if "type=RNA" in ln else 0)
        top = from_strands(seqs, nt_types=types, circular=circ)
    else:
        raise ValueError("Invalid oxDNA topology format")
    return (top, fmt) if return_format else top
