"""Development helper: host-side profile of one steady-state DiffTRe step (device-resident frames, cold pair lists).
usage: python tools/profile_step.py [frames]"""
import cProfile, pstats, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bench
from mythos_b200.energy import dna2, functional
from mythos_b200.optimization import objective
from mythos_b200.rigid_body import Quaternion
from mythos_b200.simulators.io import SimulatorTrajectory

F = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda:0")
system, c_np, q_np, obs_np = bench.make_workload(F, 0, F)
efn = dna2.create_default_energy_fn(system.topology)
theta = {k: torch.as_tensor(v, dtype=torch.float64) for k, v in efn.opt_params().items()}
kT = float(dna2.default_configs()[0]["kT"])
c_dev, q_dev = torch.from_numpy(c_np).to(dev), torch.from_numpy(q_np).to(dev)
obs = torch.tensor(obs_np, device=dev)
beta = torch.full((F,), 1.0 / kT, dtype=torch.float64, device=dev)
temperature = torch.full((F,), kT, dtype=torch.float64, device=dev)
states = SimulatorTrajectory(center=c_dev, orientation=Quaternion(q_dev), temperature=temperature, shard=(0, F, F))

def loss_fn(ref_states, weights, energy_fn, opt_params, observables):
    m = (weights * obs).sum()
    return m, (("obs", m), None)

functional.PAIR_LIST_CACHE_GB = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
with torch.no_grad():
    e_ref = objective.sharded_map(efn, states).detach()

def step():
    (l, aux), grads = objective.compute_loss_and_grad(theta, efn, beta, loss_fn, states, e_ref, [])
    host = torch.stack([grads[k] for k in sorted(grads)]).cpu()
    return float(l), host

for _ in range(6):
    step()
torch.cuda.synchronize()
ts = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(); step(); e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
    ts.append((1e3 * (t1 - t0), e0.elapsed_time(e1)))
print("wall ms / event ms per step:", [f"{a:.2f}/{b:.2f}" for a, b in ts])
# kernel-only time of one step: sum over a torch profiler trace would need kineto; use a sync'ed run of the map alone
with torch.no_grad():
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); objective.sharded_map(efn, states); e1.record(); torch.cuda.synchronize()
    print("energies-only map (no grad) ms:", e0.elapsed_time(e1))
prof = cProfile.Profile(); prof.enable(); step(); prof.disable()
pstats.Stats(prof, stream=sys.stdout).sort_stats("tottime").print_stats(28)
pstats.Stats(prof, stream=sys.stdout).sort_stats("cumulative").print_stats(45)

# GPU timeline of one step (kineto): kernels in launch order with start offsets
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as p:
    step(); torch.cuda.synchronize()
evs = [e for e in p.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
print("GPU timeline (us from first kernel): name start dur")
busy = 0.0
for e in evs:
    busy += e.time_range.end - e.time_range.start
    print(f"{e.name[:70]:70s} {e.time_range.start - t0:9.1f} {e.time_range.end - e.time_range.start:8.1f}")
print("span us", evs[-1].time_range.end - t0, "busy us", busy)
cpu = [e for e in p.events() if e.device_type == torch.autograd.DeviceType.CPU and e.name.startswith(("cudaLaunch", "cudaMemcpy", "cudaStreamSync", "cudaDeviceSync", "cudaEventSync"))]
print("host API calls:", len(cpu))
