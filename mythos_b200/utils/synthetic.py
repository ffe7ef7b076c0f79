"""Seeded synthetic systems for the benchmark configurations.

Building block: an ideal 60-bp B-form duplex (N = 120, two strands, 3'->5' internal order) from the standard
oxDNA helix construction (``ideal_duplex``).  (SURVEY 8d proposed the reference's relaxed duplex file; that
molecule is bent, so copies on an origami-like lattice interpenetrate and the excluded-volume energies become
~1e13 -- useless for parity work -- and the file would not exist on the GPU box anyway.)  ``assembly(n_dup, ...)``
places copies of the z-aligned duplex on a square lattice (pitch 2.6 oxDNA length units,
origami-like) stacked end to end in z, applies a random rotation about z and a per-nucleotide thermal jitter
(centre N(0, 0.02^2); orientation a small random rotation N(0, 0.05^2 rad)) -- the jitter keeps synthetic
configurations off the measure-zero points where acos'(+-1) is singular.
"""

from __future__ import annotations

import dataclasses as dc
import numpy as np

from mythos_b200.input.topology import Topology, bonded_pairs

WC = np.array([3, 2, 1, 0], dtype=np.int32)  # A<->T, C<->G


@dc.dataclass
class SyntheticSystem:
    center: np.ndarray  # (N,3) float64
    quat: np.ndarray  # (N,4) float64 (w,x,y,z)
    topology: Topology
    box: tuple[float, float, float]


def _quat_mul(a, b):
    w1, x1, y1, z1 = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    w2, x2, y2, z2 = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack(
        [w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
         w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2], -1)


def _rot_to_quat(axis, angle):
    axis = axis / np.linalg.norm(axis, axis=-1, keepdims=True)
    return np.concatenate([np.cos(0.5 * angle)[..., None], np.sin(0.5 * angle)[..., None] * axis], -1)


def _rotate(q, v):
    """Rotate vectors v (...,3) by unit quaternions q (...,4)."""
    w, u = q[..., :1], q[..., 1:]
    t = 2.0 * np.cross(u, v)
    return v + w * t + np.cross(u, t)


def axes_to_quaternion(a1: np.ndarray, a3: np.ndarray) -> np.ndarray:
    """(a1, a3) body axes -> unit quaternion (w,x,y,z) with a1 = R(q) e_x, a3 = R(q) e_z.

    Rotation-matrix route with the largest-component pivot (well conditioned for every rotation, including the
    half turns of the antiparallel strand); the sign of q is irrelevant, the axes are quadratic in q."""
    a1 = np.asarray(a1, dtype=np.float64)
    a3 = np.asarray(a3, dtype=np.float64)
    a2 = np.cross(a3, a1)
    m = np.stack([a1, a2, a3], axis=-1)  # columns are the body axes
    m00, m11, m22 = m[..., 0, 0], m[..., 1, 1], m[..., 2, 2]
    cand = np.stack(
        [
            np.stack([1 + m00 + m11 + m22, m[..., 2, 1] - m[..., 1, 2], m[..., 0, 2] - m[..., 2, 0], m[..., 1, 0] - m[..., 0, 1]], -1),
            np.stack([m[..., 2, 1] - m[..., 1, 2], 1 + m00 - m11 - m22, m[..., 0, 1] + m[..., 1, 0], m[..., 0, 2] + m[..., 2, 0]], -1),
            np.stack([m[..., 0, 2] - m[..., 2, 0], m[..., 0, 1] + m[..., 1, 0], 1 - m00 + m11 - m22, m[..., 1, 2] + m[..., 2, 1]], -1),
            np.stack([m[..., 1, 0] - m[..., 0, 1], m[..., 0, 2] + m[..., 2, 0], m[..., 1, 2] + m[..., 2, 1], 1 - m00 - m11 + m22], -1),
        ],
        axis=-2,
    )  # (..., 4 pivots, 4 components); row k is 4*q_k*q
    diag = np.stack([cand[..., k, k] for k in range(4)], -1)
    pick = np.argmax(diag, axis=-1)
    q = np.take_along_axis(cand, pick[..., None, None], axis=-2)[..., 0, :]
    return q / np.linalg.norm(q, axis=-1, keepdims=True)


def ideal_duplex(n_bp: int = 60, twist_deg: float = 35.9, rise: float = 0.3897628551303122, com_to_axis: float = 0.6):
    """Ideal B-form duplex along z, the standard oxDNA construction: base pair k sits at height k*rise, its a1
    rotated by k*twist about the axis; strand 1 has a3 = +z, strand 2 is the antiparallel complement (a1 and a3
    negated, listed in reverse so both strands run 3'->5' in the internal order).  -> centre (2n,3), quat (2n,4)."""
    k = np.arange(n_bp)
    ang = np.deg2rad(twist_deg) * k
    a1 = np.stack([np.cos(ang), np.sin(ang), np.zeros(n_bp)], -1)
    axis = np.stack([np.zeros(n_bp), np.zeros(n_bp), rise * k], -1)
    zhat = np.broadcast_to(np.array([0.0, 0.0, 1.0]), a1.shape)
    c = np.concatenate([axis - com_to_axis * a1, (axis + com_to_axis * a1)[::-1]])
    A1 = np.concatenate([a1, -a1[::-1]])
    A3 = np.concatenate([zhat, -zhat])
    c = c - c.mean(0)
    return c, axes_to_quaternion(A1, A3), np.array([n_bp, n_bp], dtype=np.int32)


def duplex60():
    c, q, counts = ideal_duplex(60)
    return c, q, None, counts


def jitter(center, quat, rng, sigma_pos=0.02, sigma_rot=0.05):
    c = center + rng.normal(0.0, sigma_pos, size=center.shape)
    axis = rng.normal(size=center.shape)
    ang = rng.normal(0.0, sigma_rot, size=center.shape[0])
    q = _quat_mul(_rot_to_quat(axis, ang), quat)
    return c, q / np.linalg.norm(q, axis=-1, keepdims=True)


def assembly(n_dup: int, pitch: float = 2.6, seed: int = 0, gap: float = 1.0, margin: float = 5.0,
             nt_pattern: tuple[tuple[int, int], ...] | None = None, max_columns: int | None = None,
             nicked: bool = False) -> SyntheticSystem:
    """``n_dup`` jittered copies of the duplex on a lattice.  ``nt_pattern`` cycles (strand1 type, strand2 type).
    ``nicked``: the second strand of every duplex is cut in the middle (two strands of 30), so the two nucleotides
    either side of the nick interact through the UNBONDED terms -- coaxial stacking is then active at scale."""
    rng = np.random.default_rng(seed)
    c0, q0, _, counts = duplex60()
    length = c0[:, 2].max() - c0[:, 2].min() + gap
    side = int(np.ceil(np.sqrt(n_dup))) if max_columns is None else max_columns
    per_layer = side * side
    centers, quats, seqs, nts, strands = [], [], [], [], []
    n1 = int(counts[0])
    for d in range(n_dup):
        layer, r = divmod(d, per_layer)
        gy, gx = divmod(r, side)
        phi = rng.uniform(0.0, 2.0 * np.pi)
        qz = _rot_to_quat(np.array([[0.0, 0.0, 1.0]]), np.array([phi]))[0]
        c = _rotate(qz[None], c0) + np.array([gx * pitch, gy * pitch, layer * length])
        q = _quat_mul(np.broadcast_to(qz, q0.shape), q0)
        c, q = jitter(c, q, rng)
        s1 = rng.integers(0, 4, size=n1).astype(np.int32)
        s2 = WC[s1][::-1].copy()  # strand 2 runs antiparallel; nucleotide k of strand 1 pairs with n-1-k of strand 2
        centers.append(c)
        quats.append(q)
        seqs += [s1, s2]
        t1, t2 = (1, 1) if nt_pattern is None else nt_pattern[d % len(nt_pattern)]
        nts += [np.full(n1, t1, np.int32), np.full(int(counts[1]), t2, np.int32)]
        n2 = int(counts[1])
        strands += [n1, n2 // 2, n2 - n2 // 2] if nicked else [n1, n2]
    center = np.concatenate(centers)
    center -= center.min(0) - margin
    box = tuple(float(x) for x in (center.max(0) + margin))
    n = center.shape[0]
    is_end = np.zeros(n, np.int32)
    bounds = np.cumsum([0, *strands])
    is_end[bounds[:-1]] = 1
    is_end[bounds[1:] - 1] = 1
    top = Topology(
        n_nucleotides=n, strand_counts=np.array(strands, np.int32), bonded_neighbors=bonded_pairs(strands),
        seq=np.concatenate(seqs), is_end=is_end, nt_type=np.concatenate(nts),
    )
    return SyntheticSystem(center=center, quat=np.concatenate(quats), topology=top, box=box)


def rejittered_frames(system: SyntheticSystem, n_frames: int, seed0: int = 1000, dtype=np.float64):
    """F frames = the base assembly re-jittered with default_rng(seed0 + k)."""
    cs = np.empty((n_frames, *system.center.shape), dtype=dtype)
    qs = np.empty((n_frames, *system.quat.shape), dtype=dtype)
    for k in range(n_frames):
        c, q = jitter(system.center, system.quat, np.random.default_rng(seed0 + k))
        cs[k], qs[k] = c, q
    return cs, qs
