"""One-off fixture generator for the two pieces of the hot path that live in jax_md==0.2.28 (not installable in this
repository's build image): the OrderedSparse neighbour list behind ``mythos/utils/neighbors.py:12-59`` and the rigid-body
``simulate.nvt_langevin`` step the reference passes as ``simulator_init`` (``mythos/simulators/jax_md/jaxmd.py:73,82-94``,
``examples/jaxmd/simulation.ipynb`` cell 9).

Run ONCE on any machine that has the reference's pinned stack (``pip install jax==0.9.2 jax_md==0.2.28 numpy``; CPU is
enough, a few seconds), from the root of a checkout of the REFERENCE so that ``mythos.utils.neighbors`` imports:

    JAX_ENABLE_X64=1 python /path/to/this/repo/oracle/gen_jaxmd_fixtures.py /path/to/this/repo/tests/golden

It writes ``jaxmd_neighbors.npz`` and ``jaxmd_langevin.npz``.  ``tests/test_jaxmd_fixtures.py`` consumes them when
they exist (CPU: the oracle; GPU: the kernels) and is skipped otherwise -- until then rows a18 / a20 of SURVEY section 8
stay "parity unpinned" (DESIGN.md section 2).  Test infrastructure only; nothing in ``mythos_b200/`` imports this.

Design of the cases (so that they pin behaviour, not random streams):
* neighbour list: random points in a periodic box and in free space, a few bonded chains, cutoff + dr_threshold as the
  reference calls it; saved: positions, bonds, box, cutoffs, ``idx`` (2, capacity), ``did_buffer_overflow``.  The pair
  SET (i<j, padding = N) is what is compared; jax_md's slot order inside the padded buffer is not part of the contract.
* Langevin: a quadratic rigid-body energy with closed-form gradient (E = k/2 |c - c0|^2 + kq/2 |q - q0|^2), kT = 0 so
  that the O-step is pure friction and the threefry stream drops out, non-zero initial momenta written into the state;
  saved: state before and after 1 and 5 steps (position, orientation, both momenta), dt, gamma, mass.  A second case at
  kT > 0 stores 20 000-step kinetic-energy averages for the statistical check.
"""

from __future__ import annotations

import dataclasses
import sys
from pathlib import Path

import numpy as np


def neighbour_cases(out: Path) -> None:
    import jax.numpy as jnp
    from jax_md import space

    from mythos.utils.neighbors import get_neighbor_list_fn

    rng = np.random.default_rng(11)
    cases = {}
    for tag, n, box in (("periodic", 300, 12.0), ("free", 200, 0.0)):
        pos = rng.uniform(0.0, box if box else 9.0, size=(n, 3))
        strands = np.array_split(np.arange(n), 6)
        bonded = np.concatenate([np.stack([s[:-1], s[1:]], 1) for s in strands]).astype(np.int32)
        disp, _ = space.periodic(box) if box else space.free()
        r_cutoff, dr = (2.0, 0.2) if box else (1.5, 0.1)
        nl_fn = get_neighbor_list_fn(bonded, n, disp, box if box else 9.0, r_cutoff=r_cutoff, dr_threshold=dr)
        nbrs = nl_fn.allocate(jnp.asarray(pos))
        moved = pos + rng.normal(0.0, 0.01, size=pos.shape)  # below dr/2: update() must NOT rebuild
        nbrs_same = nbrs.update(jnp.asarray(moved))
        far = pos + rng.normal(0.0, 0.3, size=pos.shape)  # beyond dr/2: update() rebuilds
        nbrs_far = nbrs.update(jnp.asarray(far))
        cases.update({
            f"{tag}_pos": pos, f"{tag}_bonded": bonded, f"{tag}_box": np.float64(box), f"{tag}_r_cutoff": np.float64(r_cutoff),
            f"{tag}_dr_threshold": np.float64(dr), f"{tag}_idx": np.asarray(nbrs.idx), f"{tag}_overflow": np.asarray(nbrs.did_buffer_overflow),
            f"{tag}_moved": moved, f"{tag}_idx_after_small_move": np.asarray(nbrs_same.idx),
            f"{tag}_far": far, f"{tag}_idx_after_large_move": np.asarray(nbrs_far.idx),
            f"{tag}_overflow_after_large_move": np.asarray(nbrs_far.did_buffer_overflow),
        })
    np.savez_compressed(out / "jaxmd_neighbors.npz", **cases)


def langevin_cases(out: Path) -> None:
    import jax
    import jax.numpy as jnp
    from jax_md import rigid_body, simulate, space

    rng = np.random.default_rng(5)
    n = 24
    c0 = rng.normal(0, 1.0, (n, 3))
    q0 = rng.normal(0, 1.0, (n, 4))
    q0 /= np.linalg.norm(q0, axis=1, keepdims=True)
    c = c0 + rng.normal(0, 0.1, (n, 3))
    q = q0 + rng.normal(0, 0.05, (n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    k_c, k_q = 3.0, 1.5

    def energy_fn(body, **kwargs):
        return 0.5 * k_c * jnp.sum((body.center - c0) ** 2) + 0.5 * k_q * jnp.sum((body.orientation.vec - q0) ** 2)

    _, shift = space.free()
    dt, kT0 = 5e-3, 296.15 * 0.1 / 300.0
    mass = rigid_body.RigidBody(jnp.array([1.0]), jnp.array([[1.0, 1.0, 1.0]]))
    gamma = rigid_body.RigidBody(jnp.array([kT0 / 2.5]), jnp.array([kT0 / 7.5]))
    body = rigid_body.RigidBody(jnp.asarray(c), rigid_body.Quaternion(jnp.asarray(q)))
    saved = {"c0": c0, "q0": q0, "k_c": k_c, "k_q": k_q, "dt": dt, "gamma_center": kT0 / 2.5, "gamma_quat": kT0 / 7.5,
             "mass": 1.0, "inertia": np.ones(3), "center": c, "quat": q}

    # deterministic case: kT = 0, momenta written by hand
    init_fn, step_fn = simulate.nvt_langevin(energy_fn, shift, dt=dt, kT=0.0, gamma=gamma)
    state = init_fn(jax.random.PRNGKey(0), body, mass=mass)
    pc = rng.normal(0, 0.3, (n, 3))
    pq_free = rng.normal(0, 0.3, (n, 4))
    pq = pq_free - (pq_free * q).sum(1, keepdims=True) * q  # conjugate momentum tangent to the unit sphere
    mom = rigid_body.RigidBody(jnp.asarray(pc), rigid_body.Quaternion(jnp.asarray(pq)))
    state = dataclasses.replace(state, momentum=mom) if dataclasses.is_dataclass(state) else state.set(momentum=mom)
    saved.update({"p_center": pc, "p_quat": pq, "force_center": np.asarray(state.force.center),
                  "force_quat": np.asarray(state.force.orientation.vec)})
    step = jax.jit(step_fn)
    for k in range(1, 6):
        state = step(state)
        if k in (1, 5):
            saved.update({f"center_{k}": np.asarray(state.position.center), f"quat_{k}": np.asarray(state.position.orientation.vec),
                          f"p_center_{k}": np.asarray(state.momentum.center), f"p_quat_{k}": np.asarray(state.momentum.orientation.vec)})

    # statistical case: kT > 0, averages over a long run (equipartition and configurational variance of the harmonic wells)
    init_fn, step_fn = simulate.nvt_langevin(energy_fn, shift, dt=dt, kT=kT0, gamma=gamma)
    state = init_fn(jax.random.PRNGKey(1), body, mass=mass)

    def body_fn(carry, _):
        st = step_fn(carry)
        ke_t = 0.5 * jnp.sum(st.momentum.center ** 2) / n
        dc2 = jnp.sum((st.position.center - c0) ** 2) / n
        return st, (ke_t, dc2)

    state, (ke_t, dc2) = jax.lax.scan(body_fn, state, jnp.arange(20000))
    saved.update({"kT": kT0, "mean_translational_ke_per_body": float(jnp.mean(ke_t[2000:])),
                  "mean_sq_displacement_per_body": float(jnp.mean(dc2[2000:]))})
    np.savez_compressed(out / "jaxmd_langevin.npz", **saved)


if __name__ == "__main__":
    import jax

    jax.config.update("jax_enable_x64", True)
    target = Path(sys.argv[1] if len(sys.argv) > 1 else "tests/golden")
    target.mkdir(parents=True, exist_ok=True)
    neighbour_cases(target)
    langevin_cases(target)
    print("wrote", target / "jaxmd_neighbors.npz", "and", target / "jaxmd_langevin.npz")
