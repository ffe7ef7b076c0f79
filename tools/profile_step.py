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

# host timeline of one step without any profiler: entry / exit stamps of the main stages (no extra syncs)
import functools
from mythos_b200.energy import theta_tape, model as kmodel
_stamps = []
def _wrap(obj, name, label=None):
    fn = getattr(obj, name)
    @functools.wraps(fn)
    def w(*a, **k):
        _stamps.append((label or name, "in", time.perf_counter()))
        try:
            return fn(*a, **k)
        finally:
            _stamps.append((label or name, "out", time.perf_counter()))
    setattr(obj, name, w)
_wrap(objective, "compute_loss"); _wrap(theta_tape, "bind"); _wrap(objective, "sharded_map"); _wrap(functional, "_run")
_wrap(functional.CellListPairs, "chunk"); _wrap(functional, "_launch"); _wrap(torch.autograd, "grad", "autograd.grad")
_wrap(functional.deferred_verification, "ok", "checks.ok"); _wrap(functional.deferred_verification, "enqueue", "checks.enqueue")
_wrap(theta_tape.FlatParams, "unflatten"); _wrap(objective, "compute_weights_and_neff")
_wrap(kmodel.Plan, "pairs", "plan.pairs"); _wrap(kmodel.Plan, "topology", "plan.topology"); _wrap(kmodel.Plan, "device_params", "plan.device_params")
_wrap(kmodel, "support_cutoffs"); _wrap(functional._FrameEnergy, "forward", "_FrameEnergy.forward"); _wrap(theta_tape.FlatParams, "__init__", "FlatParams()")
for _ in range(3):
    _stamps.clear(); torch.cuda.synchronize(); t0 = time.perf_counter(); step(); t1 = time.perf_counter()
print(f"host timeline of one step ({1e3 * (t1 - t0):.2f} ms):")
for name, io, t in _stamps:
    print(f"  {1e6 * (t - t0):8.0f} us  {io:3s} {name}")

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
