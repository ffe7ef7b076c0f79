"""Development helper: DiffTRe device pass (neighbour build + frame kernel per chunk) at several chunk sizes."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from mythos_b200.energy import dna2, functional, model as kmodel
from mythos_b200.utils import synthetic

F = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
dev = torch.device("cuda:0")
s = synthetic.assembly(17, seed=1)
c, q = synthetic.rejittered_frames(s, F)
cd, qd = torch.tensor(c, device=dev), torch.tensor(q, device=dev)
efn = dna2.create_default_energy_fn(s.topology)
plan = kmodel.plan_for(efn.energy_fns)
topo = plan.topology(cd.shape[1], dev)
params = plan.device_params(dev, torch.float64)
ones = torch.ones((F, 8), device=dev, dtype=torch.float64)
for chunk in [int(x) for x in (sys.argv[2:] or ["148", "296", "592", "1184"])]:
    functional.FRAME_CHUNK = chunk
    src = plan.pairs(dev, topo)
    fn = lambda: functional.energy_and_gradients(plan.model, topo, cd, qd, params, src, cot=ones, want_pos_grad=False, want_param_grad=True, per_frame_param_grad=True)
    fn(); fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = float(np.median(ts))
    print(f"chunk {chunk:5d}: {t:8.3f} ms per {F} frames -> {F / t * 1e3:,.0f} frames/s")
