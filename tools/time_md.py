"""Development helper: the BASELINE configs[1] MD leg alone (60-bp duplex, oxDNA1, Langevin); argv[1] = steps, argv[2] = graph|eager."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bench

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
if len(sys.argv) > 2 and sys.argv[2] == "eager":
    from mythos_b200.simulators import md
    md.MDSimulator.use_cuda_graph = False
out = bench.md_benchmark(torch.device("cuda:0"), steps)
print(f"{out['us_per_step']:.2f} us/step  {out['value']:.3e} nucleotide-steps/s  E/nt {out['final_energy_per_nt']:.4f}")
