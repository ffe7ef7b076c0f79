"""Development helper: where does the end-to-end DiffTRe step (pinned host frames, all theta) spend its time?"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from mythos_b200.energy import dna2
from mythos_b200.optimization import objective
from mythos_b200.rigid_body import Quaternion
from mythos_b200.simulators.io import SimulatorTrajectory
from mythos_b200.utils import synthetic

F = int(sys.argv[1]) if len(sys.argv) > 1 else 4736
dev = torch.device("cuda:0")
s = synthetic.assembly(17, seed=1)
c, q = synthetic.rejittered_frames(s, 64)
c = np.tile(c, (F // 64, 1, 1)); q = np.tile(q, (F // 64, 1, 1))
ch, qh = torch.from_numpy(c).pin_memory(), torch.from_numpy(q).pin_memory()
efn = dna2.create_default_energy_fn(s.topology)
theta = {k: torch.as_tensor(v, dtype=torch.float64) for k, v in efn.opt_params().items()}
kT = float(dna2.default_configs()[0]["kT"])
beta = torch.full((F,), 1.0 / kT, dtype=torch.float64, device=dev)
temperature = torch.full((F,), kT, dtype=torch.float64, device=dev)
obs = torch.randn(F, device=dev, dtype=torch.float64)
def loss_fn(ref_states, weights, energy_fn, opt_params, observables):
    m = (weights * obs).sum()
    return m, (("obs", m), None)
for host in (False, True):
    cc, qq = (ch, qh) if host else (ch.to(dev), qh.to(dev))
    states = SimulatorTrajectory(center=cc, orientation=Quaternion(qq), temperature=temperature)
    with torch.no_grad():
        e_ref = efn.map(states)
    def step():
        (l, aux), grads = objective.compute_loss_and_grad(theta, efn, beta, loss_fn, states, e_ref, [])
        return float(l), torch.stack([grads[k] for k in sorted(grads)]).cpu()
    step(); step(); torch.cuda.synchronize()
    ts = []
    for _ in range(4):
        t0 = time.perf_counter(); step(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print(f"F={F} {'pinned host' if host else 'device'} frames: {np.median(ts):.2f} ms per e2e step")
