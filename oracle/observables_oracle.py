"""CPU restatement of the reference's structural observables.  TEST INFRASTRUCTURE ONLY (imported by tests/ and smoke()).

Plain numpy, function by function as the reference writes them (paths relative to the reference repo):
  single_propeller_twist_rad      mythos/observables/propeller.py:18-31
  local_helical_axis_with_norm    mythos/observables/base.py:24-38
  single_rise                     mythos/observables/rise.py:19-37
  single_pitch_angle              mythos/observables/pitch.py:32-60
  single_diameter                 mythos/observables/diameter.py:21-43
  per-state means                 propeller.py:66-71, rise.py:66-70, pitch.py:83-88, diameter.py:71-76
PINNED by the reference's own known-answer tests (tests/test_observables_oracle.py reproduces
mythos/observables/tests/test_propeller.py:12-31,42-76, test_rise.py:14-32, test_pitch.py:38-52, test_diameter.py).
`displacement_fn(a, b)` follows jax_md: a - b, wrapped for a periodic box.
"""

from __future__ import annotations

import numpy as np

ANGSTROMS_PER_OXDNA_LENGTH = 8.518  # mythos/utils/units.py:5


def clamp(x):
    return np.clip(x, -1.0, 1.0)  # mythos/utils/math.py:78-81


def make_displacement(box=None):
    if not box:
        return lambda a, b: a - b
    L = np.asarray(box, dtype=np.float64)
    return lambda a, b: np.mod(a - b + 0.5 * L, L) - 0.5 * L


def single_propeller_twist_rad(bp, base_normals):
    return np.arccos(clamp(np.dot(base_normals[bp[0]], base_normals[bp[1]])))


def local_helical_axis_with_norm(quartet, base_sites, displacement_fn):
    (a1, b1), (a2, b2) = quartet
    midp_a1b1 = (base_sites[a1] + base_sites[b1]) / 2.0
    midp_a2b2 = (base_sites[a2] + base_sites[b2]) / 2.0
    dr = displacement_fn(midp_a2b2, midp_a1b1)
    norm = np.linalg.norm(dr)
    return dr / norm, norm


def single_rise(quartet, base_sites, displacement_fn):
    (a1, b1), (a2, b2) = quartet
    axis, _ = local_helical_axis_with_norm(quartet, base_sites, displacement_fn)
    midp1 = (base_sites[a1] + base_sites[b1]) / 2.0
    midp2 = (base_sites[a2] + base_sites[b2]) / 2.0
    dr = displacement_fn(midp2, midp1)
    return np.dot(dr, axis) * ANGSTROMS_PER_OXDNA_LENGTH


def single_pitch_angle(quartet, base_sites, back_sites, displacement_fn):
    (a1, b1), (a2, b2) = quartet
    axis, _ = local_helical_axis_with_norm(quartet, base_sites, displacement_fn)
    bb1 = displacement_fn(back_sites[b1], back_sites[a1])
    bb2 = displacement_fn(back_sites[b2], back_sites[a2])
    p1 = displacement_fn(bb1, np.dot(axis, bb1) * axis)
    p2 = displacement_fn(bb2, np.dot(axis, bb2) * axis)
    return np.arccos(clamp(np.dot(p1 / np.linalg.norm(p1), p2 / np.linalg.norm(p2))))


def single_diameter(bp, back_sites, displacement_fn, sigma_backbone):
    dr = displacement_fn(back_sites[bp[0]], back_sites[bp[1]])
    return (np.linalg.norm(dr) + sigma_backbone) * ANGSTROMS_PER_OXDNA_LENGTH


def frame_columns(base_sites, back_sites, base_normals, base_pairs, quartets, sigma_backbone, box=None):
    """(F, 4) = per-state means (propeller deg, rise A, pitch angle rad, diameter A) from (F,N,3) site arrays."""
    disp = make_displacement(box)
    F = base_sites.shape[0]
    out = np.full((F, 4), np.nan)
    bps = [] if base_pairs is None else np.asarray(base_pairs).reshape(-1, 2)
    qts = [] if quartets is None else np.asarray(quartets).reshape(-1, 2, 2)
    for f in range(F):
        if len(bps):
            out[f, 0] = np.mean([180.0 - single_propeller_twist_rad(bp, base_normals[f]) * 180.0 / np.pi for bp in bps])
            out[f, 3] = np.mean([single_diameter(bp, back_sites[f], disp, sigma_backbone) for bp in bps])
        if len(qts):
            out[f, 1] = np.mean([single_rise(q, base_sites[f], disp) for q in qts])
            out[f, 2] = np.mean([single_pitch_angle(q, base_sites[f], back_sites[f], disp) for q in qts])
    return out
