#!/usr/bin/env python
"""bench.py -- DiffTRe reweighting throughput (frames/s, E + dE/dtheta) on 1..8 B200, BASELINE.json configs[3].

Workload (``config.workload``): oxDNA2 + Debye, synthetic 17-duplex assembly (N = 2040 nt), F = 8192 stored frames
(the base assembly re-jittered per frame, seeded), float64, all-pairs semantics realised by per-frame device cell
lists at the interaction range.  One step = one full DiffTRe pass: per-frame energies AND per-frame dE/dparams rows
in one fused pass over the pair lists, Boltzmann weights / n_eff, loss = <O>_w, dL/dparams = g @ J; with N > 1 the
frames are sharded contiguously over the ranks (total fixed -> "strong"), energies all-gathered and the gradient
all-reduced over NCCL.

  value   frames/s with frames resident in HBM and the packed parameter bank on the device (CUDA events, max over ranks)
  e2e     frames/s through the public API ``compute_loss_and_grad`` with HOST (pinned) frame buffers: H2D of the frames
          (streamed chunk by chunk, overlapped with the kernels), theta -> bank chain on the host, kernels, D2H of loss
          and gradients inside the timed region
  roofline  dominant kernel (k_frame_energy: all terms of a frame, E + dE/dparams, pair lists streamed from the per-frame
          device neighbour build): algorithmic flop-equivalents / CUDA-event time vs the FP64 FMA issue peak measured in
          this run by the library's micro-benchmark (and HBM bytes vs measured copy BW)
  md / forces_8k / forces_100k   the other BASELINE configs that fit one GPU (configs[1], [2], [4]), N = 1 only
  cpu_baseline  the CPU oracle (port of the reference algorithm; the reference itself needs JAX, absent here) on a
          bounded sample of the same frames

``--impl reference`` times that oracle alone (rank 0 only).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_DUPLEX, PITCH, SEED = 17, 2.6, 1
METRIC, UNIT = "DiffTRe frames/s (E+dE/dtheta)", "frames/s"
# hardware counters of k_frame_energy<double,1,1,0> at 1184 frames x 2040 nt from the committed `ncu --set full` capture
NCU_COUNTERS = {"source": "profiles/r02_packed_k_frame_energy_details.csv", "sm__inst_executed_pipe_fp64_pct": 26.50,
                "smsp__issue_active_pct": 40.68, "sm__warps_active_pct": 24.99, "dram_bytes": 0.5614e9, "duration_ms": 2.390,
                "local_ld_st_inst": 11.1e6}
# Issue cost of the special functions in FP64 FMA slots.  MEASURED (csrc/peaks.cu micro-kernels: a dependent chain of
# f(x)*a+b against a chain of FMAs, profiles/r02_special_weights.json); bench re-measures them in every run and uses the
# live numbers -- these are only the fallback.  (SURVEY 8d's provisional guesses were div 16 sqrt 16 exp 40 log 50 acos 70.)
W_SPECIAL = {"div": 11.5, "sqrt": 12.2, "exp": 16.1, "log": 43.9, "acos": 27.5}
# SURVEY 8d per-unit work: (flops, specials) of one unit of each kind.  STRICT accounting (round 2): a term is charged only
# for the pairs on which it can be non-zero -- excluded volume per pair with a SITE pair inside its cutoff (1 short-range
# pair in 100), hydrogen bonding per pair inside its radial AND six angular supports, cross stacking per pair inside its
# radial and three plain angular supports; the other short-range pairs are charged the site screen they need (two
# nucleotides' axes and sites, six squared distances).  Work a kernel skips because it is identically zero is not credited.
WORK = {
    "nucleotide": (60, {}),
    "bonded": (290, {"acos": 3, "exp": 1, "log": 2, "sqrt": 7, "div": 3}),
    "screen": (30, {}),                                                   # centre-distance test of a listed pair
    "debye": (65, {"exp": 1, "sqrt": 1}),                                 # backbone sites inside r_cut (one reciprocal square root)
    "sr_screen": (150, {}),                                               # site screen of a pair inside the short-range centre cutoff
    "exc": (200, {"div": 4, "sqrt": 2}),                                 # a site pair inside its excluded-volume cutoff
    "hb": (165, {"acos": 6, "exp": 1, "sqrt": 1}),                        # inside the hydrogen-bond radial + six angular supports
    "cross": (165, {"acos": 6, "sqrt": 1}),                               # inside the cross-stacking radial + three angular supports
    "coax": (225, {"acos": 2, "sqrt": 1, "div": 1}),                     # stacking-site window
}


def slot_cost(kind: str) -> float:
    flop, specials = WORK[kind]
    return flop / 2 + sum(n * W_SPECIAL[k] for k, n in specials.items())


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=8192)
    ap.add_argument("--cpu-frames", type=int, default=6, help="frames in the CPU-oracle sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-legs", action="store_true", help="skip the md / forces_8k / forces_100k legs (profiling runs)")
    return ap.parse_args()


def make_workload(n_frames: int, lo: int, hi: int):
    from mythos_b200.utils import synthetic

    system = synthetic.assembly(N_DUPLEX, pitch=PITCH, seed=SEED)
    n = system.center.shape[0]
    c = np.empty((hi - lo, n, 3))
    q = np.empty((hi - lo, n, 4))
    for k in range(lo, hi):
        c[k - lo], q[k - lo] = synthetic.jitter(system.center, system.quat, np.random.default_rng(1000 + k))
    obs = np.random.default_rng(7).standard_normal(n_frames)
    return system, c, q, obs


def oracle_pass(system, c, q, frames: int):
    """CPU oracle: E and dE/d(kernel params) of `frames` frames with per-frame neighbour lists (port of the reference
    algorithm; torch autograd plays the role of jax.value_and_grad)."""
    from oracle import oxdna_oracle as orc

    top = system.topology
    theta = orc.default_theta("dna2")
    t0 = time.perf_counter()
    for f in range(frames):
        params = orc.init_all("dna2", theta)
        leaves = []
        for term in params.values():
            for k, v in list(term.items()):
                if isinstance(v, torch.Tensor) and v.dtype == torch.float64:
                    term[k] = v.detach().clone().requires_grad_(True)
                    leaves.append(term[k])
        pairs = orc.neighbor_pairs(c[f], top.bonded_neighbors, 3.25, 0.0)
        e = orc.energy_terms("dna2", c[f], q[f], top.seq, top.bonded_neighbors, pairs, params, is_end=top.is_end).sum()
        torch.autograd.grad(e, leaves, allow_unused=True)
    return frames / (time.perf_counter() - t0)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = args.cpu_frames
    system, c, q, _ = make_workload(args.frames, 0, sample)
    torch.set_num_threads(os.cpu_count() or 1)
    for _ in range(max(args.warmup, 1)):
        oracle_pass(system, c, q, 1)
    rates = [oracle_pass(system, c, q, sample) for _ in range(args.steps)]
    v = float(np.mean(rates))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sample / v, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.frames), "l2": "inputs larger than L2",
                   "sample": f"{sample} frames of the workload per step, rate extrapolated (a full 8192-frame pass of the CPU port would take ~15 min)"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{sample} frames of the same workload per step (oracle: torch-f64 restatement of the reference; JAX is not installable here)"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_name(frames: int) -> str:
    return (f"configs[3]: DiffTRe reweighting, oxDNA2+Debye, N={N_DUPLEX * 120} nt synthetic assembly "
            f"({N_DUPLEX} duplexes, pitch {PITCH}), F={frames} frames, frame-sharded, float64")


def md_benchmark(dev, n_steps: int = 10000):
    """BASELINE.json configs[1]: oxDNA1 rigid-body Langevin MD of a 60-bp duplex (N = 120), 10^4 steps, all-pairs list
    (as the reference example), float64, CUDA-graph replay of the fused step -> nucleotide-steps/s (device time)."""
    from mythos_b200 import space
    from mythos_b200.energy import dna1
    from mythos_b200.input.topology import from_strands
    from mythos_b200.rigid_body import Quaternion, RigidBody
    from mythos_b200.simulators import md
    from mythos_b200.utils import synthetic

    c, q, _ = synthetic.ideal_duplex(60)
    rng = np.random.default_rng(0)
    seq1 = "".join("ACGT"[k] for k in rng.integers(0, 4, 60))
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    top = from_strands([seq1, "".join(comp[b] for b in reversed(seq1))])
    c, q = synthetic.jitter(c, q, rng)
    efn = dna1.create_default_energy_fn(top)
    kT = 296.15 * 0.1 / 300.0
    params = md.StaticSimulatorParams(
        seq=top.seq, mass=RigidBody(torch.tensor(1.0), torch.tensor([1.0, 1.0, 1.0])),
        gamma=RigidBody(torch.tensor(kT / 2.5), torch.tensor([kT / 7.5] * 3)), bonded_neighbors=top.bonded_neighbors,
        checkpoint_every=0, dt=5e-3, kT=kT)
    sim = md.MDSimulator(energy_fn=efn, simulator_params=params, space=space.free())
    body = RigidBody(torch.tensor(c, device=dev), Quaternion(torch.tensor(q, device=dev)))
    sim.run({}, body, 200, key=1)  # warm-up (allocator, graph machinery)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    traj = sim.run({}, body, n_steps, key=2)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    e_last = float(efn(RigidBody(traj.center[-1], Quaternion(traj.orientation.vec[-1]))))
    us_step = 1e3 * ms / n_steps

    # neighbour-list variant: the same run with a cell-list neighbour list at the model's interaction range
    # (+ 0.2 skin) that updates itself on the device -- displacement test and conditional rebuild are the third launch of
    # the captured step, no host round trip (utils.neighbors.NeighborListFns on the frame-resident build)
    md_nl = None
    try:
        from mythos_b200.energy import model as kmodel
        from mythos_b200.utils import neighbors as nbmod

        rng_cut = kmodel.interaction_range(kmodel.plan_for(efn.energy_fns))
        fns = nbmod.get_neighbor_list_fn(top.bonded_neighbors, top.n_nucleotides, space.free()[0], None, r_cutoff=rng_cut, dr_threshold=0.2)
        sim_nl = md.MDSimulator(energy_fn=efn, simulator_params=params, space=space.free(), neighbors=fns)
        sim_nl.run({}, body, 200, key=1)
        torch.cuda.synchronize(dev)
        e0.record()
        traj_nl = sim_nl.run({}, body, n_steps, key=2)
        e1.record()
        torch.cuda.synchronize(dev)
        ms_nl = e0.elapsed_time(e1)
        probe = fns.allocate(body)
        for k in range(0, n_steps, 10):  # rebuild frequency along the produced trajectory (every 10th step replayed)
            probe.update(traj_nl.center[k])
        md_nl = {"us_per_step": 1e3 * ms_nl / n_steps, "value": 120 * n_steps / (ms_nl * 1e-3), "unit": "nucleotide-steps/s",
                 "r_cutoff": rng_cut, "dr_threshold": 0.2, "list_entries": int(probe.idx.shape[-1]), "pairs_listed": int(probe.count.item()),
                 "device_side_update": probe.slots is not None,
                 "rebuilds_seen_replaying_every_10th_step": int(probe.rebuilds.item()) if probe.rebuilds is not None else None,
                 "max_abs_deviation_from_all_pairs_run_first_100_steps": float((traj_nl.center[:100] - traj.center[:100]).abs().max()),
                 "launches_per_step": 3}
    except Exception as err:  # noqa: BLE001
        md_nl = {"unavailable": repr(err)[:200]}

    # Roofline of this leg: N = 120 is LAUNCH / DEPENDENCY bound, not FP or HBM bound (each step is two dependent kernels of
    # one wave: 280 B x 120 of state, ~1e5 flop).  The floor measured here is the same replay structure with empty work:
    # CUDA graphs of 16 steps x 2 dependent one-thread kernels, i.e. what the device needs just to sequence the step's two
    # launches.  frac = floor / achieved.
    x = torch.zeros(1, device=dev)
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        x.add_(1.0)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for _ in range(32):
                x.add_(1.0)
    torch.cuda.current_stream(dev).wait_stream(side)
    g.replay()
    torch.cuda.synchronize(dev)
    e0.record()
    for _ in range(256):
        g.replay()
    e1.record()
    torch.cuda.synchronize(dev)
    floor_us = 1e3 * e0.elapsed_time(e1) / (256 * 16)
    n_pairs = 7021
    flops_step = 2 * 3.0 * (slot_cost("nucleotide") * 120 + slot_cost("bonded") * 118 + slot_cost("screen") * n_pairs) + 150 * 120
    roofline = {"bound": "launch latency (two dependent kernels per step, one wave each)",
                "achieved": us_step, "peak": floor_us, "unit": "us/step", "frac": floor_us / us_step,
                "floor": "CUDA-graph replay of 2 dependent empty kernels per step, measured in this run",
                "fp64_tflops_equiv": flops_step / (us_step * 1e-6) / 1e12,
                "hbm_gbs": (35 * 8 * 120 + 8 * n_pairs) / (us_step * 1e-6) / 1e9}

    # CPU baseline of this leg: the oracle's BAOAB step (oracle/langevin_oracle.py) with forces from torch autograd over the
    # oracle energy, same system, a bounded number of steps on the host cores
    cpu = None
    try:
        from oracle import langevin_oracle as lo
        from oracle import oxdna_oracle as orc

        torch.set_num_threads(os.cpu_count() or 1)
        params_o = orc.init_all("dna1", orc.default_theta("dna1"))
        pairs_o = orc.all_unbonded_pairs(120, top.bonded_neighbors)
        cc, qq = np.array(c), np.array(q)
        pc, pq = np.zeros_like(cc), np.zeros_like(qq)
        rng_o = np.random.default_rng(0)

        def forces(cc, qq):
            ct, qt = torch.tensor(cc, requires_grad=True), torch.tensor(qq, requires_grad=True)
            e = orc.energy_terms("dna1", ct, qt, top.seq, top.bonded_neighbors, pairs_o, params_o).sum()
            gc, gq = torch.autograd.grad(e, [ct, qt])
            return gc.numpy(), gq.numpy()

        dcen, dquat = forces(cc, qq)
        n_cpu = 10
        t0 = time.perf_counter()
        for _ in range(n_cpu):
            noise = rng_o.standard_normal((120, 6))
            cc, qq, pc, pq = lo.step(cc, qq, pc, pq, dcen, dquat, noise, 5e-3, kT, kT / 2.5, kT / 7.5, 1.0, np.ones(3))[:4]
            dcen, dquat = forces(cc, qq)
        dt_cpu = time.perf_counter() - t0
        cpu = {"value": 120 * n_cpu / dt_cpu, "unit": "nucleotide-steps/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"{n_cpu} steps of the same system (oracle BAOAB step + torch-autograd forces of the oracle energy)"}
    except Exception as err:  # noqa: BLE001  (the baseline is a report, never a reason to lose the measurement)
        cpu = {"unavailable": repr(err)[:200]}
    return {"metric": "MD nucleotide-steps/s", "value": 120 * n_steps / (ms * 1e-3), "unit": "nucleotide-steps/s",
            "workload": "configs[1]: oxDNA1 Langevin MD, 60-bp duplex (N=120), 10^4 steps, all-pairs list (U=7021), float64",
            "ms_total": ms, "us_per_step": us_step, "final_energy_per_nt": e_last / 120, "roofline": roofline, "cpu_baseline": cpu,
            "neighbour_list_variant": md_nl}


# Work model: SURVEY 8d's per-pair counts, split per term so that a term counts for a pair only
# inside its radial support (flops / 2 + weighted specials = FMA-issue slots, float64):
#   distance screen of a listed pair                         30 flop
#   Debye-Hueckel (backbone sites inside r_cut)              65 flop + exp + div + sqrt
#   excluded volume (centres inside the short-range cutoff)  200 flop + 4 div + 2 sqrt
#   hydrogen bonding + cross stacking (base-site window)     330 flop + 6 acos + exp + sqrt + div
#   coaxial stacking (stacking-site window)                  225 flop + 2 acos + sqrt + div
# (sum over a pair inside every support = SURVEY's 820 flop + 8 acos + 2 exp + 5 sqrt + 7 div)


def pair_support_counts(plan, center, quat, pairs, count):
    """Per-frame means of: listed pairs, pairs inside the Debye support, the short-range centre cutoff, the
    hydrogen-bond / cross-stacking window and the coaxial window -- measured on the device from (F,2,cap) lists."""
    from mythos_b200 import _lib

    names = {nm: k for k, nm in enumerate(_lib.param_names())}
    v = plan.params_vector().detach().double()
    P = lambda nm: float(v[names[nm]])  # noqa: E731  (bank 0)
    g = plan.model.geom[0]
    F, n = center.shape[0], center.shape[1]
    i, j = pairs[:, 0].long(), pairs[:, 1].long()
    valid = (i < n) & (torch.arange(pairs.shape[-1], device=pairs.device)[None, :] < count[:, None])
    i, j = i.clamp(max=n - 1), j.clamp(max=n - 1)
    q0, q1, q2, q3 = quat.unbind(-1)
    a1 = torch.stack([q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3, 2 * (q1 * q2 + q0 * q3), 2 * (q1 * q3 - q0 * q2)], -1)
    a2 = torch.stack([2 * (q1 * q2 - q0 * q3), q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3, 2 * (q2 * q3 + q0 * q1)], -1)
    a3 = torch.stack([2 * (q1 * q3 + q0 * q2), 2 * (q2 * q3 - q0 * q1), q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3], -1)
    back = center + g.back[0] * a1 + g.back[1] * a2 + g.back[2] * a3
    base, stack = center + g.base * a1, center + g.stack * a1

    def d2(x):
        gi = torch.gather(x, 1, i.unsqueeze(-1).expand(-1, -1, 3))
        gj = torch.gather(x, 1, j.unsqueeze(-1).expand(-1, -1, 3))
        return (gi - gj).square().sum(-1)

    ob = sum(x * x for x in g.back) ** 0.5
    sr = max(P("unbonded_excluded_volume.dr_c_backbone") + 2 * ob, P("unbonded_excluded_volume.dr_c_base") + 2 * abs(g.base),
             max(P("unbonded_excluded_volume.dr_c_back_base"), P("unbonded_excluded_volume.dr_c_base_back")) + ob + abs(g.base),
             P("hydrogen_bonding.dr_c_high_hb") + 2 * abs(g.base), P("cross_stacking.dr_c_high_cross") + 2 * abs(g.base),
             P("coaxial_stacking.dr_c_high_coax") + 2 * abs(g.stack))
    in_sr = valid & (d2(center) < sr * sr)
    db = d2(back)
    in_db = valid & (db < P("debye.r_cut") ** 2) if plan.model.forms[0].has_debye else valid & False
    bb = d2(base)
    lo = min(P("hydrogen_bonding.dr_c_low_hb"), P("cross_stacking.dr_c_low_cross"))
    hi = max(P("hydrogen_bonding.dr_c_high_hb"), P("cross_stacking.dr_c_high_cross"))
    in_bp = in_sr & (bb > lo * lo) & (bb < hi * hi)
    # strict supports of the three sparse short-range terms
    def gat(x):
        return (torch.gather(x, 1, i.unsqueeze(-1).expand(-1, -1, 3)), torch.gather(x, 1, j.unsqueeze(-1).expand(-1, -1, 3)))

    (back_i, back_j), (base_i, base_j) = gat(back), gat(base)
    rc = lambda nm: P("unbonded_excluded_volume." + nm) ** 2  # noqa: E731
    in_ev = in_sr & ((db < rc("dr_c_backbone")) | (bb < rc("dr_c_base")) | ((back_i - base_j).square().sum(-1) < rc("dr_c_back_base")) |
                     ((base_i - back_j).square().sum(-1) < rc("dr_c_base_back")))
    (a1i, a1j), (a3i, a3j) = gat(a1), gat(a3)
    dh = (base_j - base_i) / bb.clamp(min=1e-30).sqrt().unsqueeze(-1)
    th = [torch.acos(x.clamp(-1, 1)) for x in (-(a1i * a1j).sum(-1), -(a1j * dh).sum(-1), (a1i * dh).sum(-1), (a3i * a3j).sum(-1),
                                               -(a3j * dh).sum(-1))] + [torch.pi - torch.acos((a3i * dh).sum(-1).clamp(-1, 1))]

    def window(term, angles):
        ok = torch.ones_like(in_sr)
        for k, t in zip(angles, th):
            t0, dc = P(f"{term}.theta0_{term_tag[term]}_{k}"), P(f"{term}.delta_theta_{term_tag[term]}_{k}_c")
            ok = ok & (t > t0 - dc) & (t < t0 + dc)
        return ok

    term_tag = {"hydrogen_bonding": "hb", "cross_stacking": "cross"}
    seq = plan.topology(n, center.device).seq.long()
    tab_w = torch.tensor([P(f"hydrogen_bonding.eps_hb_weights[{a},{b}]") if f"hydrogen_bonding.eps_hb_weights[{a},{b}]" in names else 1.0
                          for a in range(4) for b in range(4)], device=center.device)
    w_ok = tab_w[(seq[i] * 4 + seq[j])] != 0
    try:
        in_hb = in_sr & w_ok & (bb > P("hydrogen_bonding.dr_c_low_hb") ** 2) & (bb < P("hydrogen_bonding.dr_c_high_hb") ** 2) & \
            window("hydrogen_bonding", ("1", "2", "3", "4", "7", "8"))
        in_cr = in_sr & (bb > P("cross_stacking.dr_c_low_cross") ** 2) & (bb < P("cross_stacking.dr_c_high_cross") ** 2) & \
            window("cross_stacking", ("1", "2", "3"))
    except KeyError:  # (parameter naming differs: fall back to the radial window for both terms)
        in_hb = in_cr = in_bp
    ss = d2(stack)
    in_cx = in_sr & (ss > P("coaxial_stacking.dr_c_low_coax") ** 2) & (ss < P("coaxial_stacking.dr_c_high_coax") ** 2)
    mean = lambda m: float(m.sum()) / F  # noqa: E731
    return {"listed": mean(valid), "debye_support": mean(in_db), "short_range": mean(in_sr), "hb_cross_window": mean(in_bp),
            "exc_site_pair_in_range": mean(in_ev), "hb_support": mean(in_hb), "cross_support": mean(in_cr), "coax_window": mean(in_cx)}


def slots_forward(n, n_b, u, screen: bool = True):
    """Forward work of one configuration in FMA-issue slots (float64) from the measured support counts `u`.  `screen`:
    charge the centre-distance test of every listed pair (done by the neighbour walk on the support-tagged route, so NOT
    credited to the frame kernel there)."""
    return (slot_cost("nucleotide") * n + slot_cost("bonded") * n_b + (slot_cost("screen") * u["listed"] if screen else 0.0) +
            slot_cost("debye") * u["debye_support"] + slot_cost("sr_screen") * u["short_range"] + slot_cost("exc") * u["exc_site_pair_in_range"] +
            slot_cost("hb") * u["hb_support"] + slot_cost("cross") * u["cross_support"] + slot_cost("coax") * u["coax_window"])


def force_benchmark(dev, n_dup: int, model: str, seed: int, peak_tflops: float, label: str, reps: int = 7):
    """BASELINE.json configs[2] / configs[4]: neighbour rebuild + energy + forces + dE/dparams of one large configuration
    (float64, free space, neighbour list at the interaction range) -> evaluations/s, device time (CUDA events)."""
    from mythos_b200.energy import dna2, functional, na1
    from mythos_b200.energy import model as kmodel
    from mythos_b200.utils import neighbors, synthetic

    pattern = ((1, 1), (2, 2), (1, 2)) if model == "na1" else None
    s = synthetic.assembly(n_dup, seed=seed, nt_pattern=pattern)
    cd = torch.tensor(s.center, device=dev).unsqueeze(0)
    qd = torch.tensor(s.quat, device=dev).unsqueeze(0)
    n = cd.shape[1]
    efn = (na1 if model == "na1" else dna2).create_default_energy_fn(s.topology)
    plan = kmodel.plan_for(efn.energy_fns)
    topo = plan.topology(n, dev)
    params = plan.device_params(dev, torch.float64)
    cut = kmodel.interaction_range(plan)
    ones = torch.ones((1, 8), device=dev, dtype=torch.float64)
    _, count, _, ws = neighbors.build_pairs(cd, topo.bonded, (0, 0, 0), cut, 0.0, 1)
    cap = (int(int(count.item()) * 1.05) + 1024) // 4 * 4

    def timed(fn):
        fn()
        torch.cuda.synchronize(dev)
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn()
            e1.record()
            torch.cuda.synchronize(dev)
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts)), out

    t_nl, (pairs, count, ov, ws) = timed(lambda: neighbors.build_pairs(cd, topo.bonded, (0, 0, 0), cut, 0.0, cap, ws))
    assert int(ov.item()) == 0
    t_en, _ = timed(lambda: functional._launch(plan.model, topo, cd, qd, params, pairs[0], 0, 0xFF, ones, True, True, True, False, count))
    t_ef, _ = timed(lambda: functional._launch(plan.model, topo, cd, qd, params, pairs[0], 0, 0xFF, ones, True, True, False, False, count))

    def both():
        p, cnt, _, _ = neighbors.build_pairs(cd, topo.bonded, (0, 0, 0), cut, 0.0, cap, ws)
        return functional._launch(plan.model, topo, cd, qd, params, p[0], 0, 0xFF, ones, True, True, True, False, cnt)

    t_all, _ = timed(both)

    # the same rebuild + evaluation captured once in a CUDA graph and replayed (fixed shapes and capacity; the overflow
    # flag is read after the timed region): what an MD loop with a neighbour list pays per step, launch latency removed
    t_graph = None
    try:
        gstream = torch.cuda.Stream(device=dev)
        gstream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(gstream):
            both()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=gstream):
                g_out = both()
        torch.cuda.current_stream(dev).wait_stream(gstream)
        t_graph, _ = timed(graph.replay)
        ref = both()
        torch.cuda.synchronize(dev)
        if not torch.allclose(g_out[1], ref[1], rtol=1e-9, atol=1e-9 * float(ref[1].abs().max())):
            t_graph = None
    except Exception:  # noqa: BLE001  (graph capture is an optimisation of the measurement, not part of the path)
        t_graph = None

    # the same evaluation with all-pairs semantics through the product's own pair source: two support-tagged builds
    # (centres at the short-range cutoff, backbone sites at the Debye cutoff) feeding the list kernels
    from mythos_b200.input.topology import AllPairs

    src = kmodel.plan_for(efn.with_props(unbonded_neighbors=AllPairs(n)).energy_fns).pairs(dev, topo)
    src.tag_for_list_kernels = True

    def tagged():
        return functional.energy_and_gradients(plan.model, topo, cd, qd, params, src, cot=ones, want_pos_grad=True, want_param_grad=True)

    t_tag, _ = timed(tagged)
    assert src.verify()
    u = pair_support_counts(plan, cd, qd, pairs, count)
    n_b = int(topo.bonded.shape[0])
    flop_eq = 2 * 3.0 * slots_forward(n, n_b, u)  # E + forces + theta-VJP = 3 x forward; 1 FMA slot = 2 flop
    achieved = flop_eq / (t_en * 1e-3) / 1e12
    return {
        "workload": label, "n_nucleotides": n, "pairs": u,
        "metric": "evaluations/s (neighbour rebuild + energy + forces + dE/dparams)",
        "value": 1e3 / min(t for t in (t_all, t_tag, t_graph) if t),
        "nucleotide_evaluations_per_s": n * 1e3 / min(t for t in (t_all, t_tag, t_graph) if t),
        "ms": {"neighbour_rebuild": t_nl, "energy_forces_dparams": t_en, "energy_forces": t_ef, "rebuild_plus_evaluation": t_all,
               "rebuild_plus_evaluation_cuda_graph_replay": t_graph, "rebuild_plus_evaluation_support_tagged_lists": t_tag},
        "roofline": {"bound": "fp64", "kernels": "k_list_debye + k_list_sr + k_pairs<bonded> (E + forces + dE/dparams)",
                     "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved / peak_tflops if peak_tflops else None,
                     "work_model": "SURVEY 8d per-pair counts split per term, each counted inside its radial support (measured), x3 for E+F+theta-VJP; NA1: supports measured with the DNA bank's windows"},
    }


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.proc, self.path = None, f"/tmp/mb_clocks_{os.getpid()}.csv"
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(index)],
                stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.f.close()
        rows = [r.split(",") for r in Path(self.path).read_text().splitlines() if r.count(",") >= 6]
        if rows:
            sm = [float(r[0]) for r in rows if r[0].strip().replace(".", "").isdigit()]
            out["sm_mhz"] = float(np.median(sm)) if sm else None
            out["sm_max_mhz"] = float(rows[0][1])
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            out["reasons"] = [n for k, n in enumerate(names) if any("Active" in r[3 + k] and "Not" not in r[3 + k] for r in rows)]
            out["power_w_max"] = max(float(r[2]) for r in rows)
        return out


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist

    from mythos_b200 import _lib
    from mythos_b200.energy import dna2, functional
    from mythos_b200.energy import model as kmodel
    from mythos_b200.optimization import objective
    from mythos_b200.rigid_body import Quaternion
    from mythos_b200.simulators.io import SimulatorTrajectory
    from mythos_b200.utils import neighbors

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    F = args.frames
    lo, hi = objective.shard_bounds(F, rank, world)
    system, c_np, q_np, obs_np = make_workload(F, lo, hi)
    n = system.center.shape[0]
    top = system.topology
    efn = dna2.create_default_energy_fn(top)
    theta = {k: torch.as_tensor(v, dtype=torch.float64) for k, v in efn.opt_params().items()}
    kT = float(dna2.default_configs()[0]["kT"])

    c_host = torch.from_numpy(c_np).pin_memory()
    q_host = torch.from_numpy(q_np).pin_memory()
    c_dev, q_dev = c_host.to(dev), q_host.to(dev)
    obs = torch.tensor(obs_np, device=dev)
    beta = torch.full((F,), 1.0 / kT, dtype=torch.float64, device=dev)
    temperature = torch.full((hi - lo,), kT, dtype=torch.float64, device=dev)

    plan = kmodel.plan_for(efn.energy_fns)
    topo = plan.topology(n, dev)
    params_dev = plan.device_params(dev, torch.float64)
    ones = torch.ones((hi - lo, _lib.N_TERMS), dtype=torch.float64, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(steps):
            out = fn()
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out

    # FP64 FMA issue peak and the special-function weights of the work model, measured now
    from mythos_b200.utils import peaks as peaks_mod

    best = max(peaks_mod.fma_peak_tflops(dev, torch.float64) for _ in range(2))
    measured_w = peaks_mod.special_weights(dev, torch.float64)["weights"]
    W_SPECIAL.update({k: float(measured_w[k]) for k in W_SPECIAL})
    # The single-GPU legs of the other configurations run FIRST, on a quiet process: MD at N = 120 is bound by launch and
    # dependency latency, and measured 4x slower after the big passes (allocator state of the 936 MB streaming buffers) or
    # after the CPU oracle (its OpenMP workers keep spinning and slow the launch thread).
    md_line = forces_8k = forces_100k = None
    if world == 1 and rank == 0 and not args.no_legs:
        md_line = md_benchmark(dev)
        forces_8k = force_benchmark(dev, 68, "dna2", 0, best,
                                    "configs[2]: oxDNA2 + Debye, synthetic 68-duplex assembly (N=8160), neighbour list, float64")
        forces_100k = force_benchmark(dev, 834, "na1", 2, best,
                                      "configs[4]: NA1 hybrid DNA/RNA, synthetic 834-duplex assembly (N=100080), neighbour rebuild + forces, float64")
    torch.cuda.empty_cache()

    # ---- one DiffTRe step = compute_loss_and_grad (the public API): theta -> parameter bank, E and dE/dparams rows of this
    # rank's frames, all-gather, weights / n_eff, loss, dL/dE, g @ J, theta chain backward, all-reduce -> loss + dL/dtheta
    def loss_fn(ref_states, weights, energy_fn, opt_params, observables):
        measured = (weights * obs).sum()
        return measured, (("obs", measured), None)

    states_dev = SimulatorTrajectory(center=c_dev, orientation=Quaternion(q_dev), temperature=temperature, shard=(lo, hi, F))
    states_host = SimulatorTrajectory(center=c_host, orientation=Quaternion(q_host), temperature=temperature, shard=(lo, hi, F))
    with torch.no_grad():
        functional.PAIR_LIST_CACHE_GB = 0.0
        e_ref = objective.sharded_map(efn, states_dev).detach()

    debug_steps = bool(os.environ.get("MB_DEBUG_STEPS"))

    def step(states):
        t0 = time.perf_counter() if debug_steps else 0.0
        (l, aux), grads = objective.compute_loss_and_grad(theta, efn, beta, loss_fn, states, e_ref, [])
        host = torch.stack([grads[k] for k in sorted(grads)]).cpu()  # D2H of the gradients (loss below): the step's result
        if debug_steps and rank == 0:
            st = torch.cuda.memory_stats(dev)
            print(f"[step] {'dev' if states.center.is_cuda else 'host'} cache={functional.PAIR_LIST_CACHE_GB} {1e3 * (time.perf_counter() - t0):.2f} ms "
                  f"cudaMalloc {st['num_device_alloc']} cudaFree {st['num_device_free']} reserved {st['reserved_bytes.all.current'] / 2**30:.2f} GB",
                  file=sys.stderr, flush=True)
        return float(l), host, aux

    def measure(states, cache_gb):
        functional.PAIR_LIST_CACHE_GB = cache_gb
        functional._PAIR_LISTS.clear()
        for _ in range(args.warmup):
            step(states)
        # warm-up is over when a pass no longer grows the caching allocator (a cudaMalloc of one more 0.7 GB pair-list
        # block costs ~90 ms once, measured; it can still happen in the third pass): at most 4 extra untimed passes
        for _ in range(4):
            before = torch.cuda.memory_stats(dev).get("num_device_alloc", 0)
            step(states)
            if torch.cuda.memory_stats(dev).get("num_device_alloc", 0) == before:
                break
        ms, out = timed(lambda: step(states), args.steps)
        return ms / args.steps, out

    sampler = ClockSampler(local) if (rank == 0 and not os.environ.get("MB_NO_CLOCK_SAMPLER")) else None
    # headline numbers: COLD passes -- every pass rebuilds the per-frame pair lists (the pair-list cache is switched off)
    ms_step, (loss, grad_host, aux) = measure(states_dev, 0.0)
    value = F / (ms_step * 1e-3)
    if os.environ.get("MB_PROFILE_E2E") and rank == 0:  # development hook: host-side profile of one end-to-end step
        import cProfile
        import pstats

        prof = cProfile.Profile()
        prof.enable()
        step(states_host)
        prof.disable()
        pstats.Stats(prof, stream=sys.stderr).sort_stats("cumulative").print_stats(45)
    elif os.environ.get("MB_PROFILE_E2E"):
        step(states_host)
    # end to end: the frames stay in pinned host memory; the public API streams them to the device chunk by chunk (copy
    # stream), overlapped with the kernels of the previous chunk; every byte crosses PCIe inside the timed region
    ms_e2e, _ = measure(states_host, 0.0)
    clocks = sampler.stop() if sampler else None
    # WARM passes: what every pass after the first costs while an optimiser keeps reweighting the same stored frames
    # (the lists are remembered per frame tensor, functional._PairListCache; reached through DiffTReObjective.calculate)
    ms_warm, _ = measure(states_dev, 24.0)
    ms_warm_e2e, _ = measure(states_host, 24.0)
    cache_bytes = functional._PAIR_LISTS.nbytes()
    functional._PAIR_LISTS.clear()
    functional.PAIR_LIST_CACHE_GB = 0.0
    neff = aux[0]
    h2d = (c_host.numel() + q_host.numel()) * 8
    d2h = (len(theta) + 1) * 8

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- dominant kernel: unbonded pairs, E + dE/dparams, timed alone with CUDA events per launch ----
    chunk = min(1184, hi - lo)  # 8 waves of 148 CTAs
    cc, qq = c_dev[:chunk].contiguous(), q_dev[:chunk].contiguous()
    rng_cut = kmodel.interaction_range(plan)
    _, probe, _, _ = neighbors.build_pairs(cc[:8], topo.bonded, tuple(plan.model.box), rng_cut, 0.0, 1)
    cap = (int(int(probe.max()) * 1.1) + 64) // 4 * 4
    pairs, count, _, _ = neighbors.build_pairs(cc, topo.bonded, tuple(plan.model.box), rng_cut, 0.0, cap)
    u = pair_support_counts(plan, cc[:16], qq[:16], pairs[:16], count[:16])
    launches, nl_launches = [], []
    src = plan.pairs(dev, topo)  # the step's own pair source: support-tagged device lists
    tagged = src.tag is not None
    tpairs, tstride, tcount = src.chunk(slice(0, chunk), cc, qq, tagged=True, slots=True)
    u["kept_by_tagged_build"] = float(src.last_valid_count.double().mean())
    # entries the frame kernel's producer reads per frame: the packed slots' valid entries, or the padded capacity
    u["list_entries_read_by_consumer"] = float(tcount.double().mean()) if tcount is not None else tstride // 2
    for _ in range(5):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        torch.cuda.synchronize(dev)
        e0.record()
        src.chunk(slice(0, chunk), cc, qq, tagged=True, slots=True)
        e1.record()
        # the hot kernel exactly as the timed step runs it: all terms, E + J rows, the chunk's device pair lists streamed
        functional._launch(plan.model, topo, cc, qq, params_dev, tpairs, tstride, _lib.ALL_TERMS, ones[:chunk], True, False, True, True,
                           tcount, _lib.FLAG_TAGGED_PAIRS if tagged else 0)
        e2.record()
        torch.cuda.synchronize(dev)
        nl_launches.append(e0.elapsed_time(e1))
        launches.append(e1.elapsed_time(e2))
    src.verify()
    k_ms = float(np.median(launches[1:]))
    n_b = int(topo.bonded.shape[0])
    nl_ms = float(np.median(nl_launches[1:]))
    # E + params-only backward = 2.5 x forward; 1 FMA slot = 2 flop.  The kernel is credited with what IT
    # computes: on the support-tagged route the centre-distance screen of the listed pairs is the neighbour walk's work.
    kernel_flop_eq = 2 * 2.5 * chunk * slots_forward(n, n_b, u, screen=False)
    kernel_achieved = kernel_flop_eq / (k_ms * 1e-3) / 1e12
    # the headline fraction is the whole step's: all algorithmic work of the pass (screen included) over the time of the
    # COLD timed step (neighbour builds, frame kernel, reweighting, theta chain, collectives), all ranks' frames
    step_flop_eq = 2 * 2.5 * F * slots_forward(n, n_b, u, screen=True)
    step_achieved = step_flop_eq / (ms_step * 1e-3) / 1e12 / world

    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    alg_bytes = chunk * (n * 7 * 8 + 8 * u["kept_by_tagged_build"] + 232 * 8 + 64)  # frame + tagged pairs in, J row + terms row out (slot padding not counted)
    ncu = NCU_COUNTERS if (chunk == 1184 and n == 2040) else None
    roofline = {
        "bound": "fp64", "kernel": "k_frame_energy<double,WP=1> (all terms of a frame; support-tagged pair lists from the device neighbour build)",
        "achieved": step_achieved, "peak": best, "unit": "TFLOP/s", "frac": step_achieved / best if best else None,
        "scope": "whole cold step per GPU (neighbour builds + frame kernel + reweighting + theta chain); frac_kernel is the dominant kernel alone",
        "achieved_kernel": kernel_achieved, "frac_kernel": kernel_achieved / best if best else None,
        # dram__bytes_read.sum + dram__bytes_write.sum of the kernel at this shape from the committed ncu capture, in bytes;
        # other shapes have no capture.  With packed slots (the default on the frame-resident neighbour route) the list holds
        # no padding: 561 MB against 541 MB algorithmic (the padded slots read 791 MB).
        "traffic": ncu["dram_bytes"] if ncu else None,
        "ncu": ncu,
        "model_vs_counter": (None if not ncu else
                             f"frac_kernel / FP64-pipe counter = {kernel_achieved / best / (ncu['sm__inst_executed_pipe_fp64_pct'] / 100):.2f}: the model "
                             "charges a special function its measured ISSUE cost in FMA slots (most of which are integer / FP32-seed / "
                             "conversion instructions outside the FP64 pipe), the counter sees only DFMA/DMUL/DADD; compare with "
                             "smsp__issue_active instead"),
        "peak_source": "measured in this run (library FMA micro-benchmark, 148x32 blocks x 256 threads)",
        "special_weights": {k: round(v, 2) for k, v in W_SPECIAL.items()},
        "special_weights_source": "measured in this run (csrc/peaks.cu: dependent chain of f(x)*a+b vs chain of FMA)",
        "kernel_ms_per_launch": k_ms, "frames_per_launch": chunk, "neighbour_build_ms_per_chunk": nl_ms,
        "share_of_step": k_ms / (k_ms + nl_ms),
        "pairs_per_frame": u,
        "slots_forward_per_frame": {"kernel": slots_forward(n, n_b, u, screen=False), "step": slots_forward(n, n_b, u, screen=True)},
        "hbm": {"algorithmic_gb_per_launch": alg_bytes / 1e9, "achieved_gbs": alg_bytes / 1e9 / (k_ms * 1e-3), "peak_gbs": hbm_peak,
                "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
        "work_model": "SURVEY 8d per-pair flop + special counts split per term; STRICT supports measured on 16 frames: excluded volume "
                      "per pair with a site pair in range, hydrogen bonding inside radial + six angular supports, cross stacking inside "
                      "radial + three angular supports, every other short-range pair its site screen (150 flop); "
                      "slots = flop/2 + sum(special x measured weight); x2.5 for E + dE/dparams",
    }

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        oracle_pass(system, c_np, q_np, 1)
        rate = oracle_pass(system, c_np, q_np, args.cpu_frames)
        cpu_baseline = {"value": rate, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                        "sample": f"{args.cpu_frames} frames of the same workload (oracle: torch-f64 restatement of the reference algorithm, autograd for dE/dparams)"}

    n_chunks = -(-(hi - lo) // functional.FRAME_CHUNK)  # per chunk: support points + 2 exclusion-table kernels + 2 x k_nl_frame + 1 frame kernel
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(F), "frames_per_gpu": hi - lo, "n_nucleotides": n,
                   "l2": "inputs larger than L2 (frames 936 MB + 0.7 GB of pair lists per 2072-frame chunk)", "n_theta": len(theta),
                   "loss": loss, "n_eff": float(neff.detach()), "grad_norm": float(grad_host.norm()),
                   "pass": "cold: the per-frame pair lists are rebuilt in every timed step (pair-list cache off)",
                   "step": "objective.compute_loss_and_grad (theta -> bank, E + dE/dparams rows, gather, weights/n_eff, loss, g @ J, theta VJP, all-reduce)"},
        "e2e": {"value": F / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        # every pass after the first while an optimiser reweights the same stored frames (DiffTReObjective.calculate)
        "warm": {"value": F / (ms_warm * 1e-3), "ms_per_step": ms_warm, "e2e_value": F / (ms_warm_e2e * 1e-3), "e2e_ms_per_step": ms_warm_e2e,
                 "unit": UNIT, "pair_list_cache_bytes_per_gpu": cache_bytes,
                 "what": "pair lists remembered per frame tensor (built once with a 1% cutoff margin); no neighbour kernels in the timed step"},
        "gpu_launches": args.steps * (n_chunks * 6 + 1),  # k_support_points + k_nl_* + k_frame_energy per chunk, k_weights per step (torch glue not counted)
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline, "md": md_line, "forces_8k": forces_8k,
        "forces_100k": forces_100k,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
