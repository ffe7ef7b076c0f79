"""Probabilistic-sequence weights: host-side reduction of ``(pseq, SequenceConstraints)`` to what the kernel needs.

``compute_seq_dep_weight`` (``mythos/energy/utils.py:45-132``) evaluates, per pair and per call, one of four cases --
both unpaired, one paired, both paired in different base pairs, both members of the same base pair -- as explicit 4x4 /
4x4x4 expectation sums.  The first three are the SAME bilinear form ``P_i^T W P_j`` of per-nucleotide marginals
(``P`` of a paired nucleotide = its base-pair-type distribution pushed through ``BP_IDXS``), because distinct nucleotides /
base pairs are independent; only the two members of one base pair are correlated, and that expectation depends on
``(bp_pseq, W)`` alone.  So the host computes, once per parameter update and differentiably (torch):

* ``marginals(pseq, sc)``            -> ``(N,4)``
* ``same_pair_weights(bp_pseq, W)``  -> ``(n_bp,2)`` (by the position of the FIRST nucleotide inside its base pair)

and the generic pair kernel evaluates the weight of every pair it visits from them (``csrc/energy_dev.cuh: pseq_weight``).
"""

from __future__ import annotations

import numpy as np
import torch

from mythos_b200.input.sequence_constraints import BP_IDXS, SequenceConstraints

# M[w] (4 base-pair types x 4 nucleotides): one-hot of the nucleotide that position w of each base-pair type holds
_M = np.zeros((2, 4, 4))
for _t, (_a, _b) in enumerate(BP_IDXS):
    _M[0, _t, _a] = 1.0
    _M[1, _t, _b] = 1.0


def marginals(pseq, sc: SequenceConstraints) -> torch.Tensor:
    """(N,4) nucleotide distribution of every position (float64, differentiable in both parts of ``pseq``)."""
    up, bp = (torch.as_tensor(x, dtype=torch.float64) for x in pseq)
    n = sc.n_nucleotides
    rows = []
    m = torch.as_tensor(_M, dtype=torch.float64, device=up.device)
    for i in range(n):
        if int(sc.is_unpaired[i]):
            rows.append(up[int(sc.idx_to_unpaired_idx[i])])
        else:
            k, w = (int(x) for x in sc.idx_to_bp_idx[i])
            rows.append(bp[k] @ m[w])
    return torch.stack(rows)


def same_pair_weights(pseq, table: torch.Tensor, sc: SequenceConstraints) -> torch.Tensor:
    """(n_bp,2): expectation of ``table`` for the pair made of the two members of base pair k, ``[k][w]`` with w the
    position of the pair's first nucleotide (``utils.py:96-104``: ``sum_t p[t] W[BP[t][w_i], BP[t][w_j]]``)."""
    bp = torch.as_tensor(pseq[1], dtype=torch.float64)
    table = torch.as_tensor(table, dtype=torch.float64)
    a, b = BP_IDXS[:, 0], BP_IDXS[:, 1]
    first0 = table[a, b]  # within_i = 0, within_j = 1
    first1 = table[b, a]
    if sc.n_bp == 0:
        return torch.zeros((0, 2), dtype=torch.float64)
    return torch.stack([bp[: sc.n_bp] @ first0, bp[: sc.n_bp] @ first1], dim=1)


def index_arrays(sc: SequenceConstraints) -> tuple[np.ndarray, np.ndarray]:
    """(bp_of (N) int32 with -1 for unpaired, within (N) int32)."""
    m = np.asarray(sc.idx_to_bp_idx, dtype=np.int32)
    return np.ascontiguousarray(m[:, 0]), np.ascontiguousarray(np.maximum(m[:, 1], 0))


def seq_dep_weight(pseq, i: int, j: int, table, sc: SequenceConstraints) -> torch.Tensor:
    """The weight of one pair the way the kernel evaluates it (host reference of ``pseq_weight``; used by tests)."""
    P = marginals(pseq, sc)
    bi, bj = int(sc.idx_to_bp_idx[i][0]), int(sc.idx_to_bp_idx[j][0])
    if bi >= 0 and bi == bj:
        return same_pair_weights(pseq, table, sc)[bi, int(sc.idx_to_bp_idx[i][1])]
    return P[i] @ torch.as_tensor(table, dtype=torch.float64) @ P[j]
