"""Development helper: times the frame-resident kernel in all-pairs (in-kernel cell list) and list mode."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from mythos_b200 import _lib
from mythos_b200.energy import dna2, functional, model as kmodel
from mythos_b200.utils import synthetic, neighbors

F = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda:0")
s = synthetic.assembly(17, seed=1)
c, q = synthetic.rejittered_frames(s, F)
cd, qd = torch.tensor(c, device=dev), torch.tensor(q, device=dev)
efn = dna2.create_default_energy_fn(s.topology)
plan = kmodel.plan_for(efn.energy_fns)
topo = plan.topology(cd.shape[1], dev)
params = plan.device_params(dev, torch.float64)
cut = kmodel.interaction_range(plan)
ones = torch.ones((F, 8), device=dev, dtype=torch.float64)

def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), out

for wp in (True, False):
    t_cell, o1 = timeit(lambda: functional._launch(plan.model, topo, cd, qd, params, None, 0, 0xFF, ones, True, False, wp, True, None, 0, cut))
    pairs, count, ov, ws = neighbors.build_pairs(cd, topo.bonded, (0, 0, 0), cut, 0.0, 100000)
    t_nl, _ = timeit(lambda: neighbors.build_pairs(cd, topo.bonded, (0, 0, 0), cut, 0.0, 100000, ws))
    t_list, o2 = timeit(lambda: functional._launch(plan.model, topo, cd, qd, params, pairs, 2 * 100000, 0xFF, ones, True, False, wp, True, count))
    K = 80
    mr = torch.zeros((F,), dtype=torch.int32, device=dev)
    rp, rcount, rov, ws = neighbors.build_pairs(cd, topo.bonded, (0, 0, 0), cut, 0.0, K * cd.shape[1], ws, rows=True, max_row=mr)
    t_rows, _ = timeit(lambda: neighbors.build_pairs(cd, topo.bonded, (0, 0, 0), cut, 0.0, K * cd.shape[1], ws, rows=True, max_row=mr))
    t_rlist, o3 = timeit(lambda: functional._launch(plan.model, topo, cd, qd, params, rp, 2 * K * cd.shape[1], 0xFF, ones, True, False, wp, True, None))
    print(f"   rows mode (K={K}, longest row {int(mr.max())}, overflow {int(rov)}): nl {t_rows:.3f} ms + frame kernel {t_rlist:.3f} ms; max|dE| {float((o3[0]-o2[0]).abs().max()):.2e}")
    print(f"WP={wp} F={F}: in-kernel all-pairs {t_cell:.3f} ms | nl build {t_nl:.3f} ms + list mode {t_list:.3f} ms; max|dE| {float((o1[0]-o2[0]).abs().max()):.2e}")
