"""CUDA path against the ORACLE at the scale of BASELINE's configs (not kernel against kernel).

Round-1's large-system checks compared one kernel family with another; here every route the large configurations
take is compared with the oracle's per-term energies, forces, dE/dquat and dE/d(kernel parameters, every bank):

* oxDNA2, N = 4320 (36 duplexes): plain neighbour list -> list kernels (k_list_debye + k_list_sr), the AllPairs
  sentinel's own route, and the support-tagged two-build route (configs[2]'s path);
* NA1 hybrid DNA/RNA, N = 4320 with the ((1,1),(2,2),(1,2)) duplex pattern -> MULTI-bank list kernels, plain and
  tagged (configs[4]'s path);
* the DiffTRe system (N = 2040): forces of stored frames (list kernels) and the frame kernel's energies + Jacobian rows.

Tolerance: 1e-6 relative (float64), the north star's bar, on arrays scaled by their largest magnitude.
"""

import numpy as np
import pytest
import torch

from mythos_b200 import _lib
from mythos_b200.energy import dna2, functional, na1
from mythos_b200.energy import model as kmodel
from mythos_b200.input.topology import AllPairs
from mythos_b200.utils import neighbors, synthetic
from oracle import oxdna_oracle as orc
from tests.test_device_math_host import _leafify, _oracle_param_grads

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _oracle_all(model, s, center, quat, cot, cutoff):
    """Oracle terms (8,), dE/dcenter, dE/dquat, dE/dparams (n_banks*P, kernel order) of sum_t cot_t E_t for one frame."""
    top = s.topology
    params = _leafify(orc.init_all(model, orc.default_theta(model)))
    c = torch.tensor(center, requires_grad=True)
    q = torch.tensor(quat, requires_grad=True)
    pairs = orc.neighbor_pairs(center, top.bonded_neighbors, cutoff, 0.0)
    t = orc.energy_terms(model, c, q, top.seq, top.bonded_neighbors, pairs, params, is_end=top.is_end,
                         nt_type=top.nt_type if model == "na1" else None)
    (t * torch.as_tensor(cot)).sum().backward()
    n_banks = 3 if model == "na1" else 1
    return t.detach().numpy(), c.grad.numpy(), q.grad.numpy(), _oracle_param_grads(model, params, n_banks), pairs.shape[1]


def _structural_zeros(want_p, pvec):
    names, P = _lib.param_names(), _lib.param_count()
    for b in range(len(pvec) // P):
        for i, nm in enumerate(names):
            if "[" in nm and pvec[b * P + i] == 0.0:  # HB_WEIGHTS_SA zeros: structural for the kernels
                want_p[b * P + i] = 0.0
    return want_p


def _close(got, want, what, rtol=1e-6):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    scale = max(float(np.abs(want).max()), 1e-300)
    np.testing.assert_allclose(got, want, rtol=rtol, atol=rtol * 1e-1 * scale, err_msg=what)


@pytest.mark.parametrize("model", ["dna2", "na1"])
def test_large_assembly_list_kernels_match_oracle(model):
    pattern = ((1, 1), (2, 2), (1, 2)) if model == "na1" else None
    s = synthetic.assembly(36, seed=4, nt_pattern=pattern, nicked=True)  # nicks: coaxial stacking is active at scale
    n = s.center.shape[0]
    assert n >= 4000
    mod = na1 if model == "na1" else dna2
    efn = mod.create_default_energy_fn(s.topology)
    plan = kmodel.plan_for(efn.energy_fns)
    cut = kmodel.interaction_range(plan)
    cot = np.random.default_rng(5).uniform(0.5, 1.5, size=8)
    want_t, want_c, want_q, want_p, n_pairs = _oracle_all(model, s, s.center, s.quat, cot, cut)
    assert n_pairs > 100_000 and np.all(want_t != 0.0)
    pvec = plan.params_vector().detach().numpy()
    want_p = _structural_zeros(want_p, pvec)

    cd = torch.tensor(s.center[None], device=DEV)
    qd = torch.tensor(s.quat[None], device=DEV)
    topo = plan.topology(n, cd.device)
    params = plan.device_params(cd.device, torch.float64)
    cotd = torch.tensor(cot[None], device=DEV)

    # (i) plain device neighbour list at the interaction range -> list kernels
    pairs, count, ov, _ = neighbors.build_pairs(cd, topo.bonded, (0, 0, 0), cut, 0.0, (n_pairs + 1024) // 4 * 4)
    assert int(ov.item()) == 0 and int(count.item()) == n_pairs  # same pair set as the oracle's O(N^2) search
    routes = {"plain list -> list kernels": functional.energy_and_gradients(
        plan.model, topo, cd, qd, params, functional.StaticPairs(pairs[0]), cot=cotd, want_pos_grad=True, want_param_grad=True)}
    # (ii) the AllPairs sentinel's own route, (iii) the support-tagged two-build route
    for tagged in (False, True):
        src = kmodel.plan_for(efn.with_props(unbonded_neighbors=AllPairs(n)).energy_fns).pairs(cd.device, topo)
        src.tag_for_list_kernels = tagged
        routes[f"AllPairs, tagged={tagged}"] = functional.energy_and_gradients(
            plan.model, topo, cd, qd, params, src, cot=cotd, want_pos_grad=True, want_param_grad=True)
    for what, (t, dc, dq, dp) in routes.items():
        _close(t[0].cpu().numpy(), want_t, f"{model} {what}: terms", rtol=1e-9)
        _close(dc[0].cpu().numpy(), want_c, f"{model} {what}: dE/dcenter")
        _close(dq[0].cpu().numpy(), want_q, f"{model} {what}: dE/dquat")
        _close(dp.cpu().numpy(), want_p, f"{model} {what}: dE/dparams")


def test_difftre_system_forces_and_jacobian_rows_match_oracle():
    """configs[3]'s system (N = 2040): forces of two stored frames, frame-kernel energies and per-frame dE/dparams rows."""
    s = synthetic.assembly(17, seed=1)
    cs, qs = synthetic.rejittered_frames(s, 2)
    n = s.center.shape[0]
    efn = dna2.create_default_energy_fn(s.topology).with_props(unbonded_neighbors=AllPairs(n))
    plan = kmodel.plan_for(efn.energy_fns)
    cut = kmodel.interaction_range(plan)
    rng = np.random.default_rng(8)
    cot = rng.uniform(0.5, 1.5, size=(2, 8))
    cd, qd = torch.tensor(cs, device=DEV), torch.tensor(qs, device=DEV)
    topo = plan.topology(n, cd.device)
    params = plan.device_params(cd.device, torch.float64)
    pvec = plan.params_vector().detach().numpy()
    cotd = torch.tensor(cot, device=DEV)
    # forces (list kernels) and per-frame Jacobian rows (frame-resident kernel, support-tagged warp-slot lists)
    t_f, dc, dq, _ = functional.energy_and_gradients(plan.model, topo, cd, qd, params, plan.pairs(cd.device, topo), cot=cotd,
                                                     want_pos_grad=True, want_param_grad=False)
    t_j, _, _, J = functional.energy_and_gradients(plan.model, topo, cd, qd, params, plan.pairs(cd.device, topo), cot=cotd,
                                                   want_pos_grad=False, want_param_grad=True, per_frame_param_grad=True)
    for f in range(2):
        want_t, want_c, want_q, want_p, _ = _oracle_all("dna2", s, cs[f], qs[f], cot[f], cut)
        want_p = _structural_zeros(want_p, pvec)
        _close(t_f[f].cpu().numpy(), want_t, "terms (force route)", rtol=1e-9)
        _close(t_j[f].cpu().numpy(), want_t, "terms (frame kernel)", rtol=1e-9)
        _close(dc[f].cpu().numpy(), want_c, "dE/dcenter")
        _close(dq[f].cpu().numpy(), want_q, "dE/dquat")
        _close(J[f].cpu().numpy(), want_p, "frame kernel dE/dparams row")
