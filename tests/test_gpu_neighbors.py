"""Neighbour build: the pair SET must equal the O(N^2) evaluation of the reference's accept test, bit for bit.

The reference only checks that bonded pairs are absent on 3 particles (mythos/utils/tests/test_neighbors.py:18-41);
here the set is compared on synthetic assemblies in free space and in periodic boxes, in both dtypes, batched, plus
capacity overflow reporting and energy equality between the listed pairs and the all-pairs list.
"""

import numpy as np
import pytest
import torch

from mythos_b200 import space
from mythos_b200.energy import dna2
from mythos_b200.rigid_body import Quaternion, RigidBody
from mythos_b200.utils import neighbors, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def brute_force(center: np.ndarray, bonded: np.ndarray, cutoff: float, box, dtype):
    """All i<j, not bonded, d2 < cutoff^2 with d2 accumulated as (dx*dx + dy*dy) + dz*dz in `dtype` (no FMA)."""
    c = center.astype(dtype)
    n = c.shape[0]
    i, j = np.triu_indices(n, k=1)
    d = c[i] - c[j]
    if box is not None:
        L = np.asarray(box, dtype=dtype)
        s = np.fmod(d + dtype(0.5) * L, L)
        s = np.where((s != 0) & (s < 0), s + L, s)
        d = s - dtype(0.5) * L
    d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    cut = dtype(cutoff)
    keep = d2 < cut * cut
    bset = {(min(a, b), max(a, b)) for a, b in bonded.tolist()}
    return {(int(a), int(b)) for a, b in zip(i[keep], j[keep]) if (int(a), int(b)) not in bset}


def to_set(pairs: torch.Tensor, n: int):
    p = pairs.cpu().numpy()
    valid = p[0] < n
    assert np.all(p[1][~valid] == n), "padding must be N in both rows"
    assert np.all(p[0][valid] < p[1][valid]), "OrderedSparse: i < j"
    out = set(zip(p[0][valid].tolist(), p[1][valid].tolist()))
    assert len(out) == int(valid.sum()), "duplicate pairs in the list"
    return out


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("periodic", [False, True])
@pytest.mark.parametrize("cutoff", [1.7, 3.45])
def test_pair_set_bit_exact(dtype, periodic, cutoff):
    s = synthetic.assembly(6, seed=3)
    npdt = np.float64 if dtype == torch.float64 else np.float32
    center = s.center.copy()
    box = None
    if periodic:
        box = (9.5, 8.0, 26.0)  # smaller than the assembly: nucleotides interact through the boundary
    frames = np.stack([center, center + 0.3, synthetic.jitter(center, s.quat, np.random.default_rng(5))[0]])
    c = torch.tensor(frames, dtype=dtype, device=DEV)
    pairs, count, overflow, _ = neighbors.build_pairs(
        c, torch.tensor(s.topology.bonded_neighbors), box or (0.0, 0.0, 0.0), cutoff - 0.2, 0.2, 60000)
    assert int(overflow.item()) == 0
    n = center.shape[0]
    for f in range(3):
        want = brute_force(np.asarray(c[f].cpu()), s.topology.bonded_neighbors, npdt(npdt(cutoff - 0.2) + npdt(0.2)), box, npdt)
        got = to_set(pairs[f], n)
        assert got == want, (len(got), len(want), list(got ^ want)[:5])
        assert int(count[f]) == len(want)


@pytest.mark.parametrize("periodic", [False, True])
def test_rows_layout_same_pair_set(periodic):
    """MB_NL_ROWS (one pass, fixed-width rows, unused slots = N) holds the same pair set as the compact list, reports the
    longest row, and flags rows that are too narrow."""
    s = synthetic.assembly(6, seed=3)
    box = (9.5, 8.0, 26.0) if periodic else (0.0, 0.0, 0.0)
    c = torch.tensor(np.stack([s.center, s.center + 0.3]), device=DEV)
    bonded = torch.tensor(s.topology.bonded_neighbors)
    n = s.center.shape[0]
    pairs, count, overflow, ws = neighbors.build_pairs(c, bonded, box, 3.0, 0.2, 60000)
    mr = torch.zeros((2,), dtype=torch.int32, device=DEV)
    rows, rcount, rov, ws = neighbors.build_pairs(c, bonded, box, 3.0, 0.2, 128 * n + 17, rows=True, max_row=mr)
    assert int(rov.item()) == 0 and torch.equal(rcount, count)
    for f in range(2):
        assert to_set(rows[f], n) == to_set(pairs[f], n)
    width = int(mr.max().item())
    assert 0 < width <= 128
    _, _, rov2, _ = neighbors.build_pairs(c, bonded, box, 3.0, 0.2, (width - 1) * n, rows=True, max_row=mr)
    assert int(rov2.item()) & 1


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("periodic", [False, True])
def test_warp_slot_layout_same_pair_set(periodic, dtype):
    """MB_NL_WARP_SLOTS (one pass: each warp of 32 cell-ordered nucleotides fills its own fixed-width slot, rest padded with
    N) holds the compact list's pair set, carries the tag bits, reports the sizes it needed and flags slots / lane rows
    that are too small; two builds share one list through slot_base."""
    s = synthetic.assembly(6, seed=3)
    box = (9.5, 8.0, 26.0) if periodic else (0.0, 0.0, 0.0)
    c = torch.tensor(np.stack([s.center, s.center + 0.3]), device=DEV, dtype=dtype)
    bonded = torch.tensor(s.topology.bonded_neighbors)
    n = s.center.shape[0]
    wpf = (n + 31) // 32
    ref, rcount, _, ws = neighbors.build_pairs(c, bonded, box, 2.0, 0.2, 60000)
    ref2, _, _, ws = neighbors.build_pairs(c, bonded, box, 1.2, 0.0, 60000, ws)
    W1, W2 = 32 * 64, 32 * 32
    cap = wpf * (W1 + W2)
    pairs = torch.empty((2, 2, cap), dtype=torch.int32, device=DEV)
    count = torch.empty((2,), dtype=torch.int32, device=DEV)
    ov = torch.zeros((1,), dtype=torch.int32, device=DEV)
    mr1, mr2 = (torch.empty((2, 2), dtype=torch.int32, device=DEV) for _ in range(2))
    neighbors.build_pairs(c, bonded, box, 2.0, 0.2, cap, ws, tag_bits=1 << 30, out=(pairs, count, ov), max_row=mr1, warp_slots=(96, 0, W1))
    neighbors.build_pairs(c, bonded, box, 1.2, 0.0, cap, ws, tag_bits=1 << 29, out=(pairs, count, ov), max_row=mr2, warp_slots=(96, wpf * W1, W2))
    assert int(ov.item()) == 0
    for f in range(2):
        p = pairs[f].cpu().numpy()
        valid = p[0] < n
        assert np.all(p[1][~valid] == n)
        tag = p[1][valid] >> 29
        idx = np.stack([p[0][valid], p[1][valid] & 0x1FFFFFFF])
        first = set(zip(idx[0][tag == 2].tolist(), idx[1][tag == 2].tolist()))
        second = set(zip(idx[0][tag == 1].tolist(), idx[1][tag == 1].tolist()))
        assert first == to_set(ref[f], n) and second == to_set(ref2[f], n)
        assert int(count[f]) == len(first) + len(second)
        # the first build's entries all sit in its own slots
        assert np.all(np.nonzero(valid)[0][tag == 2] < wpf * W1)
    lane_max, warp_max = (int(x) for x in mr1.max(0).values.tolist())
    assert 0 < lane_max <= 96 and lane_max <= warp_max <= W1
    _, _, ov2, _ = neighbors.build_pairs(c, bonded, box, 2.0, 0.2, wpf * (warp_max - 1), ws, max_row=mr1, warp_slots=(96, 0, warp_max - 1))
    assert int(ov2.item()) & 1
    _, _, ov3, _ = neighbors.build_pairs(c, bonded, box, 2.0, 0.2, cap, ws, max_row=mr1, warp_slots=(lane_max - 1, 0, W1))
    assert int(ov3.item()) & 4


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("n_duplex,cutoff", [(17, 2.1), (17, 1.0), (3, 3.4), (40, 1.6)])
def test_frame_resident_build_writes_the_multi_launch_lists(dtype, n_duplex, cutoff, monkeypatch):
    """Free-space warp-slot builds run as ONE launch (k_nl_frame: cell table and records of a frame in shared memory); the
    multi-launch route (MYTHOS_B200_NL_FRAME=0) is the referee: lists, counts, slot statistics and flags are identical."""
    s = synthetic.assembly(n_duplex, seed=11)
    rng = np.random.default_rng(2)
    frames = np.stack([s.center] + [synthetic.jitter(s.center, s.quat, rng)[0] for _ in range(6)])
    frames[3] += 40.0  # a frame far from the origin
    c = torch.tensor(frames, dtype=dtype, device=DEV)
    bonded = torch.tensor(s.topology.bonded_neighbors)
    n, F = s.center.shape[0], frames.shape[0]
    wpf = (n + 31) // 32
    out = {}
    for route in ("0", "1"):
        monkeypatch.setenv("MYTHOS_B200_NL_FRAME", route)
        for lane_slots, W in ((96, 32 * 48), (9, 200)):  # roomy, and too small on purpose (overflow flags, truncation)
            cap = wpf * W + 8
            pairs = torch.full((F, 2, cap), -7, dtype=torch.int32, device=DEV)
            count = torch.empty((F,), dtype=torch.int32, device=DEV)
            ov = torch.zeros((1,), dtype=torch.int32, device=DEV)
            mr = torch.empty((F, 2), dtype=torch.int32, device=DEV)
            neighbors.build_pairs(c, bonded, (0.0, 0.0, 0.0), cutoff, 0.1, cap, None, tag_bits=1 << 30, out=(pairs, count, ov),
                                  max_row=mr, warp_slots=(lane_slots, 0, W))
            out[(route, lane_slots)] = (pairs.cpu(), count.cpu(), ov.cpu(), mr.cpu())
    for lane_slots in (96, 9):
        a, b = out[("0", lane_slots)], out[("1", lane_slots)]
        for x, y, what in zip(a, b, ("pairs", "count", "overflow", "max_row")):
            assert torch.equal(x, y), (what, lane_slots)
    assert int(out[("1", 96)][2].item()) == 0 and int(out[("1", 96)][1].min()) > 0
    want = brute_force(np.asarray(c[4].cpu()), s.topology.bonded_neighbors,
                       (np.float64 if dtype == torch.float64 else np.float32)(cutoff) + (np.float64 if dtype == torch.float64 else np.float32)(0.1),
                       None, np.float64 if dtype == torch.float64 else np.float32) if n <= 2100 else None
    if want is not None:
        p = out[("1", 96)][0][4].numpy()
        valid = (p[0] < n) & (p[0] >= 0)  # (the 8 entries behind the last slot are never written)
        got = set(zip(p[0][valid].tolist(), (p[1][valid] & 0x1FFFFFFF).tolist()))
        assert got == want


def test_capacity_overflow_is_reported_and_list_truncated():
    s = synthetic.assembly(2, seed=1)
    c = torch.tensor(s.center[None], device=DEV)
    pairs, count, overflow, _ = neighbors.build_pairs(c, torch.tensor(s.topology.bonded_neighbors), (0, 0, 0), 3.0, 0.2, 100)
    assert int(overflow.item()) & 1 and int(count[0]) > 100
    p = pairs[0].cpu().numpy()
    assert np.all(p[0] < p[1]) and p.max() < s.center.shape[0]


def test_listed_pairs_give_all_pairs_energy_and_update_protocol(monkeypatch):
    monkeypatch.setattr(neighbors, "DEVICE_SIDE_UPDATE", False)  # the reference's control flow: host-checked update, new object
    s = synthetic.assembly(2, seed=4)
    top = s.topology
    efn = dna2.create_default_energy_fn(top)  # N=240 < 512: explicit all-pairs list
    body = RigidBody(torch.tensor(s.center, device=DEV), Quaternion(torch.tensor(s.quat, device=DEV)))
    e_all = efn.compute_terms(body)
    fns = neighbors.get_neighbor_list_fn(top.bonded_neighbors, top.n_nucleotides, space.free()[0], None, r_cutoff=3.3, dr_threshold=0.2)
    nbrs = fns.allocate(body)
    assert nbrs.idx.shape[0] == 2 and int(nbrs.did_buffer_overflow.item()) == 0
    e_nl = efn.with_props(unbonded_neighbors=nbrs.idx).compute_terms(body)
    np.testing.assert_allclose(e_nl.cpu().numpy(), e_all.cpu().numpy(), rtol=1e-12, atol=1e-12)
    # small move: no rebuild (same object back); large move: rebuilt list, still exact
    same = nbrs.update(body.center + 0.01)
    assert same is nbrs
    moved = RigidBody(body.center + torch.tensor([0.5, 0.0, 0.0], device=DEV) * (torch.arange(top.n_nucleotides, device=DEV) % 2).unsqueeze(1), body.orientation)
    nbrs2 = nbrs.update(moved.center)
    assert nbrs2 is not nbrs
    e_all2 = efn.compute_terms(moved)
    e_nl2 = efn.with_props(unbonded_neighbors=nbrs2.idx).compute_terms(moved)
    if int(nbrs2.did_buffer_overflow.item()) == 0:
        np.testing.assert_allclose(e_nl2.cpu().numpy(), e_all2.cpu().numpy(), rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_packed_slots_are_the_padded_slots_without_the_padding(dtype):
    """MB_NL_PACKED_SLOTS (frame-resident route): two tagged builds into one list, the warps' slots back to back -- exactly the
    entries of the padded layout in the same order, count[f] of them at the head; too small a capacity is flagged."""
    s = synthetic.assembly(17, seed=11)
    rng = np.random.default_rng(3)
    frames = np.stack([synthetic.jitter(s.center, s.quat, rng)[0] for _ in range(5)])
    c = torch.tensor(frames, dtype=dtype, device=DEV)
    bonded = torch.tensor(s.topology.bonded_neighbors)
    n, F = s.center.shape[0], frames.shape[0]
    wpf = (n + 31) // 32
    W1, W2, K = 400, 900, 80
    cap = wpf * (W1 + W2)

    def build(packed):
        pairs = torch.full((F, 2, cap), -7, dtype=torch.int32, device=DEV)
        count = torch.empty((F,), dtype=torch.int32, device=DEV)
        ov = torch.zeros((1,), dtype=torch.int32, device=DEV)
        mr1, mr2 = (torch.empty((F, 2), dtype=torch.int32, device=DEV) for _ in range(2))
        _, _, _, ws = neighbors.build_pairs(c, bonded, (0.0, 0.0, 0.0), 1.7, 0.0, cap, None, tag_bits=1 << 30, out=(pairs, count, ov), max_row=mr1,
                                            warp_slots=(K, 0, W1), packed_slots=packed)
        neighbors.build_pairs(c, bonded, (0.0, 0.0, 0.0), 2.4, 0.0, cap, ws, tag_bits=1 << 29, out=(pairs, count, ov), max_row=mr2,
                              warp_slots=(K, wpf * W1, W2), packed_slots=packed, reuse_exclusions=True)
        return pairs.cpu().numpy(), count.cpu().numpy(), int(ov.item()), mr1.cpu(), mr2.cpu()

    pad, pad_count, pad_ov, a1, a2 = build(False)
    pk, pk_count, pk_ov, b1, b2 = build(True)
    assert pad_ov == 0 and pk_ov == 0 and np.array_equal(pad_count, pk_count) and torch.equal(a1, b1) and torch.equal(a2, b2)
    for f in range(F):
        valid = pad[f, 0] < n
        assert int(valid.sum()) == int(pk_count[f])
        assert np.array_equal(pk[f][:, : pk_count[f]], pad[f][:, valid])
        assert np.all(pk[f][:, pk_count[f]:] == -7)  # nothing is written behind the packed entries
    # capacity one entry short of what frame 0 needs: flagged, count clamped
    small = int(pk_count.max()) - 1
    pairs = torch.empty((F, 2, small), dtype=torch.int32, device=DEV)
    count = torch.empty((F,), dtype=torch.int32, device=DEV)
    ov = torch.zeros((1,), dtype=torch.int32, device=DEV)
    _, _, _, ws = neighbors.build_pairs(c, bonded, (0.0, 0.0, 0.0), 1.7, 0.0, small, None, tag_bits=1 << 30, out=(pairs, count, ov),
                                        warp_slots=(K, 0, W1), packed_slots=True)
    neighbors.build_pairs(c, bonded, (0.0, 0.0, 0.0), 2.4, 0.0, small, ws, tag_bits=1 << 29, out=(pairs, count, ov),
                          warp_slots=(K, 1, W2), packed_slots=True)
    assert int(ov.item()) & 1 and int(count.max()) <= small


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_device_side_update_rebuilds_in_place_only_after_a_move(dtype):
    """Free-space lists of small systems live in the warp-slot layout and update() is ONE conditional launch: no rebuild
    below dr_threshold/2 (list, reference and counter untouched), a rebuild in place above it (reference <- positions,
    counter + 1, energies of the listed pairs = energies of all pairs); a slot that became too small is flagged."""
    s = synthetic.assembly(2, seed=4)
    top = s.topology
    efn = dna2.create_default_energy_fn(top)
    body = RigidBody(torch.tensor(s.center, device=DEV, dtype=dtype), Quaternion(torch.tensor(s.quat, device=DEV, dtype=dtype)))
    tol = 1e-12 if dtype == torch.float64 else 2e-4
    e_all = efn.compute_terms(body)
    fns = neighbors.get_neighbor_list_fn(top.bonded_neighbors, top.n_nucleotides, space.free()[0], None, r_cutoff=3.3, dr_threshold=0.2)
    nbrs = fns.allocate(body)
    assert nbrs.slots is not None and nbrs.idx.shape[0] == 2 and int(nbrs.did_buffer_overflow.item()) == 0
    assert nbrs.reference_position.shape == body.center.shape
    n = top.n_nucleotides
    want = brute_force(np.asarray(body.center.cpu()), top.bonded_neighbors, np.asarray(body.center.cpu()).dtype.type(3.3) + np.asarray(body.center.cpu()).dtype.type(0.2),
                       None, np.asarray(body.center.cpu()).dtype.type)
    assert to_set(nbrs.idx, n) == want and int(nbrs.count.item()) == len(want)
    e_nl = efn.with_props(unbonded_neighbors=nbrs.idx).compute_terms(body)
    np.testing.assert_allclose(e_nl.cpu().numpy(), e_all.cpu().numpy(), rtol=tol, atol=tol)
    before, ref_before = nbrs.idx.clone(), nbrs.reference_position.clone()
    # 0.09 < dr_threshold / 2: nothing happens
    same = nbrs.update(body.center + torch.tensor([0.09, 0.0, 0.0], device=DEV, dtype=dtype))
    assert same is nbrs and int(nbrs.rebuilds.item()) == 0
    assert torch.equal(nbrs.idx, before) and torch.equal(nbrs.reference_position, ref_before)
    # every other nucleotide moves by 0.5: rebuilt in place
    moved = RigidBody(body.center + torch.tensor([0.5, 0.0, 0.0], device=DEV, dtype=dtype)
                      * (torch.arange(n, device=DEV) % 2).unsqueeze(1).to(dtype), body.orientation)
    again = nbrs.update(moved.center)
    assert again is nbrs and int(nbrs.rebuilds.item()) == 1
    assert torch.equal(nbrs.reference_position, moved.center)
    if int(nbrs.did_buffer_overflow.item()) == 0:
        e_nl2 = efn.with_props(unbonded_neighbors=nbrs.idx).compute_terms(moved)
        np.testing.assert_allclose(e_nl2.cpu().numpy(), efn.compute_terms(moved).cpu().numpy(), rtol=tol, atol=tol)
    nbrs.update(moved.center)  # same positions again: no second rebuild
    assert int(nbrs.rebuilds.item()) == 1
    nbrs.update(moved.center, force_rebuild=True)
    assert int(nbrs.did_buffer_overflow.item()) in (0, 1, 4, 5)
    # a collapse of the structure overflows the slots: reported, not silently truncated
    squeezed = moved.center * 0.3
    nbrs.update(squeezed)
    assert int(nbrs.did_buffer_overflow.item()) & 5


def test_conditional_rebuild_decides_frame_by_frame():
    """mb_nl_args.reference with several frames: every frame is tested against its own reference positions; only the
    frames that moved are rebuilt (list, count, reference, counter), the others are left exactly as they were."""
    s = synthetic.assembly(3, seed=6)
    bonded = torch.tensor(s.topology.bonded_neighbors)
    n = s.center.shape[0]
    wpf = (n + 31) // 32
    rng = np.random.default_rng(8)
    frames = np.stack([synthetic.jitter(s.center, s.quat, rng)[0] for _ in range(4)])
    c = torch.tensor(frames, device=DEV)
    K, W = 64, 32 * 40
    cap = wpf * W
    pairs = torch.empty((4, 2, cap), dtype=torch.int32, device=DEV)
    count = torch.empty((4,), dtype=torch.int32, device=DEV)
    ov = torch.zeros((1,), dtype=torch.int32, device=DEV)
    mr = torch.empty((4, 2), dtype=torch.int32, device=DEV)
    _, _, _, ws = neighbors.build_pairs(c, bonded, (0.0, 0.0, 0.0), 2.0, 0.2, cap, None, out=(pairs, count, ov), max_row=mr, warp_slots=(K, 0, W))
    ref = c.clone()
    rebuilds = torch.zeros((4,), dtype=torch.int32, device=DEV)
    before, count_before = pairs.clone(), count.clone()
    moved = c.clone()
    moved[1] += torch.tensor([0.3, 0.0, 0.0], device=DEV) * (torch.arange(n, device=DEV) % 3 == 0).unsqueeze(1)  # beyond 0.1
    moved[3, 5, 2] += 0.11                                                                                      # one nucleotide, just beyond
    moved[2] += 0.05                                                                                            # rigid shift below the threshold
    neighbors.build_pairs(moved, bonded, (0.0, 0.0, 0.0), 2.0, 0.2, cap, ws, out=(pairs, count, ov), max_row=mr, warp_slots=(K, 0, W),
                          reference=ref, move_threshold=0.1, rebuilds=rebuilds, reuse_exclusions=True)
    assert rebuilds.tolist() == [0, 1, 0, 1] and int(ov.item()) == 0
    for f in (0, 2):
        assert torch.equal(pairs[f], before[f]) and torch.equal(ref[f], c[f]) and int(count[f]) == int(count_before[f])
    for f in (1, 3):
        assert torch.equal(ref[f], moved[f])
        want = brute_force(np.asarray(moved[f].cpu()), s.topology.bonded_neighbors, np.float64(2.0) + np.float64(0.2), None, np.float64)
        assert to_set(pairs[f], n) == want and int(count[f]) == len(want)
    # conditional rebuilds are refused where the frame-resident route does not apply (periodic box)
    with pytest.raises(Exception):
        neighbors.build_pairs(moved, bonded, (30.0, 30.0, 60.0), 2.0, 0.2, cap, ws, out=(pairs, count, ov), max_row=mr, warp_slots=(K, 0, W),
                              reference=ref, move_threshold=0.1, rebuilds=rebuilds)


def test_invalid_bonded_index_raises():
    with pytest.raises(ValueError):
        neighbors.get_neighbor_list_fn(np.array([[0, 5]]), 3, space.free()[0], None)
