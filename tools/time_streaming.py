"""Development helper: does the chunked host->device streaming of map() overlap with the kernels?"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from mythos_b200.energy import dna2
from mythos_b200.rigid_body import Quaternion, RigidBody
from mythos_b200.utils import synthetic

F = int(sys.argv[1]) if len(sys.argv) > 1 else 4736
dev = torch.device("cuda:0")
s = synthetic.assembly(17, seed=1)
c, q = synthetic.rejittered_frames(s, 64)
c = np.tile(c, (F // 64, 1, 1)); q = np.tile(q, (F // 64, 1, 1))
ch, qh = torch.from_numpy(c).pin_memory(), torch.from_numpy(q).pin_memory()
cd, qd = ch.to(dev), qh.to(dev)
efn = dna2.create_default_energy_fn(s.topology)
th = {"eps_hb": torch.tensor(1.07, dtype=torch.float64, requires_grad=True)}

def run(cc, qq):
    e = efn.with_params(th).map(RigidBody(cc, Quaternion(qq)))
    e.sum().backward()
    return e

def timeit(fn, n=4):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts))

print(f"F={F}: device-resident {timeit(lambda: run(cd, qd)):.2f} ms | pinned host, streamed {timeit(lambda: run(ch, qh)):.2f} ms | "
      f"plain H2D of all frames {timeit(lambda: (ch.to(dev, non_blocking=True), qh.to(dev, non_blocking=True))):.2f} ms "
      f"({(ch.numel() + qh.numel()) * 8 / 1e6:.0f} MB)")
