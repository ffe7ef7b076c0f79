#!/usr/bin/env python
"""bench.py -- DiffTRe reweighting throughput (frames/s, E + dE/dtheta) on 1..8 B200, BASELINE.json configs[3].

Workload (``config.workload``): oxDNA2 + Debye, synthetic 17-duplex assembly (N = 2040 nt), F = 8192 stored frames
(the base assembly re-jittered per frame, seeded), float64, all-pairs semantics realised by per-frame device cell
lists at the interaction range.  One step = one full DiffTRe pass: per-frame energies AND per-frame dE/dparams rows
in one fused pass over the pair lists, Boltzmann weights / n_eff, loss = <O>_w, dL/dparams = g @ J; with N > 1 the
frames are sharded contiguously over the ranks (total fixed -> "strong"), energies all-gathered and the gradient
all-reduced over NCCL.

  value   frames/s with frames resident in HBM and the packed parameter bank on the device (CUDA events, max over ranks)
  e2e     frames/s through the public API ``compute_loss_and_grad`` with HOST (pinned) frame buffers: H2D of the frames,
          theta -> bank chain on the host, kernels, D2H of loss and gradients inside the timed region
  roofline  dominant kernel (unbonded pairs, E + dE/dparams): algorithmic flop-equivalents / CUDA-event time vs the
          FP64 FMA issue peak measured in this run by the library's micro-benchmark (and HBM bytes vs measured copy BW)
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
# SURVEY 8d provisional per-unit work (flops, FMA-equivalent special slots) for float64
W_SPECIAL = {"div": 16, "sqrt": 16, "exp": 40, "log": 50, "acos": 70}
S_BONDED = 3 * 70 + 40 + 2 * 50 + 7 * 16 + 3 * 16
S_LR = 40 + 16 + 16
S_SR = 8 * 70 + 2 * 40 + 5 * 16 + 7 * 16


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=8192)
    ap.add_argument("--cpu-frames", type=int, default=6, help="frames in the CPU-oracle sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
        "config": {"workload": workload_name(args.frames), "l2": "inputs larger than L2"},
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
    return {"metric": "MD nucleotide-steps/s", "value": 120 * n_steps / (ms * 1e-3), "unit": "nucleotide-steps/s",
            "workload": "configs[1]: oxDNA1 Langevin MD, 60-bp duplex (N=120), 10^4 steps, all-pairs list (U=7021), float64",
            "ms_total": ms, "us_per_step": 1e3 * ms / n_steps, "final_energy_per_nt": e_last / 120}


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
    source = plan.pairs(dev, topo)
    ones = torch.ones((hi - lo, _lib.N_TERMS), dtype=torch.float64, device=dev)

    def gather(e_local):
        return e_local if world == 1 else objective.gather_frames(e_local, F)

    def device_pass(e_ref):
        terms, _, _, J = functional.energy_and_gradients(
            plan.model, topo, c_dev, q_dev, params_dev, source, cot=ones, want_pos_grad=False, want_param_grad=True,
            per_frame_param_grad=True)
        e = gather(terms.sum(1)).requires_grad_(True)
        w, neff = objective.compute_weights_and_neff(beta, e, e_ref)
        loss = (w * obs).sum()
        (g,) = torch.autograd.grad(loss, e)
        dp = g[lo:hi] @ J
        if world > 1:
            dist.all_reduce(dp)
        return loss, neff, dp, e.detach()

    with torch.no_grad():
        t0, _, _, _ = functional.energy_and_gradients(plan.model, topo, c_dev, q_dev, params_dev, source, want_pos_grad=False)
        e_ref = gather(t0.sum(1))

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

    for _ in range(args.warmup):
        device_pass(e_ref)
    sampler = ClockSampler(local) if rank == 0 else None
    ms_total, (loss, neff, dp, _) = timed(lambda: device_pass(e_ref), args.steps)
    ms_step = ms_total / args.steps
    value = F / (ms_step * 1e-3)

    # ---- end to end through the public API, host buffers ----
    def loss_fn(ref_states, weights, energy_fn, opt_params, observables):
        measured = (weights * obs).sum()
        return measured, (("obs", measured), None)

    def e2e_step():
        cd = c_host.to(dev, non_blocking=True)
        qd = q_host.to(dev, non_blocking=True)
        states = SimulatorTrajectory(center=cd, orientation=Quaternion(qd), temperature=temperature, shard=(lo, hi, F))
        (l, aux), grads = objective.compute_loss_and_grad(theta, efn, beta, loss_fn, states, e_ref, [])
        host = torch.stack([grads[k] for k in sorted(grads)]).cpu()
        return float(l), host

    for _ in range(max(1, args.warmup - 1)):
        e2e_step()
    ms_e2e, _ = timed(e2e_step, args.steps)
    ms_e2e /= args.steps
    clocks = sampler.stop() if sampler else None
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
    cap = int(int(probe.max()) * 1.1) + 64
    pairs, count, _, _ = neighbors.build_pairs(cc, topo.bonded, tuple(plan.model.box), rng_cut, 0.0, cap)
    u_nl = float(count.double().mean())
    i, j = pairs[:8, 0].long(), pairs[:8, 1].long()
    valid = i < n
    d = torch.gather(cc[:8], 1, i.clamp(max=n - 1).unsqueeze(-1).expand(-1, -1, 3)) - torch.gather(cc[:8], 1, j.clamp(max=n - 1).unsqueeze(-1).expand(-1, -1, 3))
    u_sr = float(((d.square().sum(-1) < 1.675**2) & valid).sum()) / 8
    u_lr = u_nl - u_sr
    launches = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        e0.record()
        # the hot kernel exactly as the timed step runs it: all terms, all-pairs mode (in-kernel cell list), E + J rows
        functional._launch(plan.model, topo, cc, qq, params_dev, None, 0, _lib.ALL_TERMS, ones[:chunk], True, False, True, True,
                           None, 0, rng_cut)
        e1.record()
        torch.cuda.synchronize(dev)
        launches.append(e0.elapsed_time(e1))
    k_ms = float(np.median(launches[1:]))
    n_b = int(topo.bonded.shape[0])
    slots_fwd = chunk * ((60 * n + 290 * n_b + 30 * u_nl + 65 * u_lr + 820 * u_sr) / 2 + S_BONDED * n_b + S_LR * u_lr + S_SR * u_sr)
    flop_eq = 2 * 2.5 * slots_fwd  # E + params-only backward = 2.5 x forward; 1 FMA slot = 2 flop
    achieved = flop_eq / (k_ms * 1e-3) / 1e12

    # FP64 FMA issue peak, measured
    scratch = torch.empty(148 * 32 * 256, dtype=torch.float64, device=dev)
    lib = _lib.lib()
    best = 0.0
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.mythos_b200_fma_peak_f64(_lib.current_stream(dev), scratch.data_ptr(), 148 * 32, 4096), "fma_peak")
        e1.record()
        torch.cuda.synchronize(dev)
        best = max(best, 148 * 32 * 256 * 4096 * 16 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    alg_bytes = chunk * (n * 7 * 8 + 232 * 8 + 64)  # frame in, J row + terms row out; pair lists never exist in HBM
    roofline = {
        "bound": "fp64", "kernel": "k_frame_energy<double,WP=1> (all terms, in-kernel cell list)", "achieved": achieved, "peak": best, "unit": "TFLOP/s",
        "frac": achieved / best if best else None, "traffic": None,
        "peak_source": "measured in this run (library FMA micro-benchmark, 148x32 blocks x 256 threads)",
        "kernel_ms_per_launch": k_ms, "frames_per_launch": chunk,
        "pairs_per_frame": {"listed": u_nl, "long_range_only": u_lr, "short_range": u_sr},
        "hbm": {"algorithmic_gb_per_launch": alg_bytes / 1e9, "achieved_gbs": alg_bytes / 1e9 / (k_ms * 1e-3), "peak_gbs": hbm_peak,
                "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
        "work_model": "SURVEY 8d provisional per-pair counts x measured pair numbers; fp64 special weights div16 sqrt16 exp40 log50 acos70",
    }

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        oracle_pass(system, c_np, q_np, 1)
        rate = oracle_pass(system, c_np, q_np, args.cpu_frames)
        cpu_baseline = {"value": rate, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                        "sample": f"{args.cpu_frames} frames of the same workload (oracle: torch-f64 restatement of the reference algorithm, autograd for dE/dparams)"}

    md_line = None
    if world == 1:
        md_line = md_benchmark(dev)

    n_chunks = 1  # all-pairs mode: one frame-kernel launch covers the rank's whole block of frames
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(F), "frames_per_gpu": hi - lo, "n_nucleotides": n,
                   "l2": "inputs larger than L2 (frames 936 MB + per-chunk pair lists)", "n_theta": len(theta),
                   "loss": float(loss.detach()), "n_eff": float(neff.detach()), "grad_norm": float(dp.norm())},
        "e2e": {"value": F / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": args.steps * (n_chunks * 1 + 1),  # k_frame_energy + k_weights per step (torch glue not counted)
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline, "md": md_line,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
