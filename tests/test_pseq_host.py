"""Probabilistic-sequence weights on the host: the marginal / same-pair reduction the kernel consumes
(mythos_b200/energy/pseq.py) against the case-by-case restatement of ``compute_seq_dep_weight``
(oracle, mythos/energy/utils.py:45-132), which in turn reproduces the reference's hand-expanded known answers
(mythos/energy/tests/test_utils.py:14-165)."""

import itertools

import numpy as np
import pytest
import torch

from mythos_b200.energy import pseq as kpseq
from mythos_b200.input import sequence_constraints as jd_sc
from oracle import oxdna_oracle as orc

UP = [[0.27, 0.03, 0.68, 0.02], [0.04, 0.56, 0.22, 0.18]]
BP = [[0.66, 0.14, 0.01, 0.19]]
W = [[0.2, 0.1, 0.3, 0.4], [0.05, 0.25, 0.1, 0.6], [0.55, 0.15, 0.2, 0.1], [0.1, 0.15, 0.6, 0.15]]
IS_UP, TO_UP, TO_BP = [0, 1, 1, 0], [-1, 0, 1, -1], [[0, 0], [-1, -1], [-1, -1], [0, 1]]


def test_reference_known_answers():
    """test_utils.py:14-60: the same-base-pair case (0,3) and the paired-unpaired case (0,1), expanded by hand there."""
    pseq = (torch.tensor(UP, dtype=torch.float64), torch.tensor(BP, dtype=torch.float64))
    same = orc.compute_seq_dep_weight(pseq, 0, 3, W, IS_UP, TO_UP, TO_BP)
    np.testing.assert_allclose(float(same), 0.66 * 0.4 + 0.14 * 0.1 + 0.01 * 0.15 + 0.19 * 0.1, rtol=1e-14)
    mixed = orc.compute_seq_dep_weight(pseq, 0, 1, W, IS_UP, TO_UP, TO_BP)
    # nt0 is position 0 of the base pair: AT -> A, TA -> T, GC -> G, CG -> C; nt1 is unpaired with distribution UP[0]
    want = sum(BP[0][t] * UP[0][b] * W[a][b] for t, a in enumerate((0, 3, 2, 1)) for b in range(4))
    np.testing.assert_allclose(float(mixed), want, rtol=1e-14)


@pytest.mark.parametrize("seed", [0, 1])
def test_marginal_formulation_equals_case_by_case(seed):
    rng = np.random.default_rng(seed)
    n = 10
    sc = jd_sc.from_bps(n, np.array([[0, 9], [2, 6], [3, 4]]))
    up = rng.random((sc.n_unpaired, 4))
    up /= up.sum(1, keepdims=True)
    bp = rng.random((sc.n_bp, 4))
    bp /= bp.sum(1, keepdims=True)
    table = rng.random((4, 4))
    pseq = (torch.tensor(up), torch.tensor(bp))
    for i, j in itertools.permutations(range(n), 2):
        want = orc.compute_seq_dep_weight(pseq, i, j, table, sc.is_unpaired, sc.idx_to_unpaired_idx, sc.idx_to_bp_idx)
        got = kpseq.seq_dep_weight(pseq, i, j, table, sc)
        np.testing.assert_allclose(float(got), float(want), rtol=1e-13, err_msg=f"pair ({i},{j})")
    np.testing.assert_allclose(kpseq.marginals(pseq, sc).sum(1).numpy(), np.ones(n), rtol=1e-14)


def test_constraints_and_one_hot_sequences():
    sc = jd_sc.from_bps(6, np.array([[0, 5], [1, 4]]))
    assert sc.n_unpaired == 2 and sc.n_bp == 2 and sc.unpaired.tolist() == [2, 3]
    assert sc.idx_to_bp_idx.tolist() == [[0, 0], [1, 0], [-1, -1], [-1, -1], [1, 1], [0, 1]]
    up, bp = jd_sc.dseq_to_pseq([0, 2, 1, 3, 1, 3], sc)  # A G C T C T: pairs A-T and G-C
    assert up.tolist() == [[0, 1, 0, 0], [0, 0, 0, 1]] and bp.tolist() == [[1, 0, 0, 0], [0, 0, 1, 0]]
    P = kpseq.marginals((up, bp), sc)
    assert P.argmax(1).tolist() == [0, 2, 1, 3, 1, 3]
    with pytest.raises(ValueError, match=jd_sc.ERR_DSEQ_TO_PSEQ_INVALID_BP):
        jd_sc.dseq_to_pseq([0, 2, 1, 3, 1, 0], sc)  # A paired with A
    with pytest.raises(ValueError, match=jd_sc.ERR_BP_ARR_CONTAINS_DUPLICATES):
        jd_sc.from_bps(6, np.array([[0, 5], [0, 4]]))
    with pytest.raises(ValueError, match=jd_sc.ERR_INVALID_BP_INDICES):
        jd_sc.from_bps(6, np.array([[0, 6]]))
    empty = jd_sc.from_bps(4, np.zeros((0, 2), dtype=np.int64))
    assert jd_sc.dseq_to_pseq([0, 1, 2, 3], empty)[1].shape == (1, 4)
