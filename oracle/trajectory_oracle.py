"""CPU restatement of the reference's trajectory parser.  TEST INFRASTRUCTURE ONLY.

Follows ``_read_file`` (mythos/input/trajectory.py:249-301: "t", "b", "E" header lines, then N lines of 15 numbers parsed
with np.fromstring; per-strand reversal when ``is_5p_3p``) and ``NucleotideState.quaternions`` (trajectory.py:163-175 ->
mythos/utils/math.py:9-65).  PINNED: ``tests/golden/*.npz`` hold centres / quaternions that ``oracle/build_fixtures.py``
produced from the reference's own ``data/test-data/*/output.dat`` files by the same formulas, and
``tests/golden/traj_dna1_simple_helix_head.dat`` is the head of one of those files (tests/test_trajectory_oracle.py)."""

from __future__ import annotations

import itertools

import numpy as np


def principal_axes_to_euler_angles(x, y, z):
    return np.arctan2(x[:, 1], x[:, 0]), np.arcsin(-np.clip(x[:, 2], -1, 1)), np.arctan2(y[:, 2], z[:, 2])


def euler_angles_to_quaternion(psi, theta, phi):
    sp, cp = np.sin(0.5 * psi), np.cos(0.5 * psi)
    st, ct = np.sin(0.5 * theta), np.cos(0.5 * theta)
    sf, cf = np.sin(0.5 * phi), np.cos(0.5 * phi)
    return np.array([sp * st * sf + cp * ct * cf, -sp * st * cf + sf * cp * ct, sp * ct * sf + cp * st * cf, sp * ct * cf - cp * st * sf]).T


def read_text(text: str, strand_lengths, is_5p_3p: bool = True):
    """-> times (F), boxes (F,3), energies (F,3), states (F,N,15)."""
    n = sum(strand_lengths)
    bounds = list(itertools.pairwise([0, *itertools.accumulate(strand_lengths)]))
    ts, bs, es, states, state = [], [], [], [], []
    for line in text.splitlines():
        if not line.strip():
            continue
        if line[0] == "t":
            ts.append(float(line.strip().split("=")[1]))
        elif line[0] == "b":
            bs.append(np.array(line.strip().split("=")[1].split(), dtype=np.float64))
        elif line[0] == "E":
            es.append(np.array(line.strip().split("=")[1].split(), dtype=np.float64))
        else:
            state.append(np.array(line.split(), dtype=np.float64))
            if len(state) == n:
                if is_5p_3p:
                    state = list(itertools.chain.from_iterable([state[s:e][::-1] for s, e in bounds]))
                states.append(np.array(state, dtype=np.float64))
                state = []
    return np.array(ts), np.array(bs), np.array(es), np.array(states)


def rigid_bodies(states: np.ndarray):
    """(F,N,15) -> centres (F,N,3), quaternions (F,N,4)."""
    quats = []
    for s in states:
        a1, a3 = s[:, 3:6], s[:, 6:9]
        quats.append(euler_angles_to_quaternion(*principal_axes_to_euler_angles(a1, np.cross(a3, a1), a3)))
    return states[:, :, :3].copy(), np.array(quats)
