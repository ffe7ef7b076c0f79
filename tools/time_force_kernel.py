"""Development helper: large-system energy + forces + dE/dparams (BASELINE configs[2] and configs[4] shapes).

Times the neighbour build and the energy call (phase-queued list kernel vs the generic one-thread-per-pair kernel),
checks that both kernels agree, and prints the pair statistics the roofline model needs.

usage: python tools/time_force_kernel.py [n_duplexes=68] [model=dna2|na1] [dtype=f64|f32]
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

from mythos_b200 import _lib
from mythos_b200.energy import dna2, functional, model as kmodel, na1
from mythos_b200.utils import neighbors, synthetic

n_dup = int(sys.argv[1]) if len(sys.argv) > 1 else 68
model = sys.argv[2] if len(sys.argv) > 2 else "dna2"
dtype = torch.float32 if (len(sys.argv) > 3 and sys.argv[3] == "f32") else torch.float64
dev = torch.device("cuda:0")

pattern = ((1, 1), (2, 2), (1, 2)) if model == "na1" else None
s = synthetic.assembly(n_dup, seed=0 if n_dup == 68 else 2, nt_pattern=pattern)
cd = torch.tensor(s.center, device=dev, dtype=dtype).unsqueeze(0)
qd = torch.tensor(s.quat, device=dev, dtype=dtype).unsqueeze(0)
N = cd.shape[1]
efn = (na1 if model == "na1" else dna2).create_default_energy_fn(s.topology)
plan = kmodel.plan_for(efn.energy_fns)
topo = plan.topology(N, dev)
params = plan.device_params(dev, dtype)
cut = kmodel.interaction_range(plan)
ones = torch.ones((1, 8), device=dev, dtype=dtype)
print(f"N={N} model={model} dtype={dtype} interaction range {cut:.4f}")


def timeit(fn, n=7):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), out


cap0 = 80 * N
pairs, count, ov, ws = neighbors.build_pairs(cd, topo.bonded, (0, 0, 0), cut, 0.0, cap0)
U = int(count.item())
assert U <= cap0 and int(ov.item()) == 0, (U, cap0, int(ov.item()))
cap = U + 1000
t_nl, (pairs, count, ov, ws) = timeit(lambda: neighbors.build_pairs(cd, topo.bonded, (0, 0, 0), cut, 0.0, cap, ws))
print(f"listed pairs {U} ({U / N:.1f} per nt); nl build {t_nl:.3f} ms")

res = {}
for name, flags in (("list", 0), ("generic", _lib.FLAG_GENERIC_KERNEL)):
    for what, (wf, wp) in (("E", (False, False)), ("E+F", (True, False)), ("E+F+dP", (True, True))):
        t, out = timeit(lambda: functional._launch(plan.model, topo, cd, qd, params, pairs[0], 0, 0xFF, ones, True, wf, wp, False, count, flags))
        res[(name, what)] = out
        print(f"{name:8s} {what:7s} {t:8.3f} ms   {N / t * 1e3:.3e} nt-evals/s")
for what in ("E", "E+F", "E+F+dP"):
    a, b = res[("list", what)], res[("generic", what)]
    msg = f"{what}: terms rel {float(((a[0] - b[0]).abs() / (b[0].abs() + 1e-30)).max()):.2e}"
    if a[1] is not None:
        msg += f" | d_center {float((a[1] - b[1]).abs().max() / b[1].abs().max()):.2e} | d_quat {float((a[2] - b[2]).abs().max() / b[2].abs().max()):.2e}"
    if a[3] is not None:
        msg += f" | d_params {float((a[3] - b[3]).abs().max() / b[3].abs().max()):.2e}"
    print(msg)
print("terms", res[("list", "E")][0].cpu().numpy().round(3))
