"""Development helper: per-pass times of cold DiffTRe passes over device-resident / pinned host frames (repeats shown)."""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
os.environ["MYTHOS_B200_DEBUG"] = "1"
import numpy as np, torch
import bench
from mythos_b200.energy import dna2, functional
from mythos_b200.optimization import objective
from mythos_b200.rigid_body import Quaternion
from mythos_b200.simulators.io import SimulatorTrajectory

F = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = torch.device("cuda:0")
system, c_np, q_np, obs_np = bench.make_workload(F, 0, F)
efn = dna2.create_default_energy_fn(system.topology)
theta = {k: torch.as_tensor(v, dtype=torch.float64) for k, v in efn.opt_params().items()}
kT = float(dna2.default_configs()[0]["kT"])
c_host, q_host = torch.from_numpy(c_np).pin_memory(), torch.from_numpy(q_np).pin_memory()
c_dev, q_dev = c_host.to(dev), q_host.to(dev)
obs = torch.tensor(obs_np, device=dev)
beta = torch.full((F,), 1.0 / kT, dtype=torch.float64, device=dev)
temperature = torch.full((F,), kT, dtype=torch.float64, device=dev)
def loss_fn(ref_states, weights, energy_fn, opt_params, observables):
    m = (weights * obs).sum(); return m, (("obs", m), None)
sd = SimulatorTrajectory(center=c_dev, orientation=Quaternion(q_dev), temperature=temperature)
sh = SimulatorTrajectory(center=c_host, orientation=Quaternion(q_host), temperature=temperature)
functional.PAIR_LIST_CACHE_GB = 0.0
with torch.no_grad():
    e_ref = efn.map(sd).detach()
for name, st in (("device", sd), ("host", sh), ("device", sd)):
    for k in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        (l, aux), g = objective.compute_loss_and_grad(theta, efn, beta, loss_fn, st, e_ref, [])
        torch.cuda.synchronize()
        ms_ = torch.cuda.memory_stats()
        print(name, k, round(1e3 * (time.perf_counter() - t0), 2), "ms", "cudaMalloc calls", ms_["num_device_alloc"], "frees", ms_["num_device_free"],
              "retries", ms_["num_alloc_retries"], "reserved GB", round(ms_["reserved_bytes.all.current"] / 2**30, 2),
              "memo", list(functional._SIZING._data.values()), flush=True)
