"""Measure the FMA issue peak and the special-function weights (FMA-slot equivalents) on the GPU; writes
gpurun_out/special_weights.json (copied to profiles/ by hand once reviewed)."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mythos_b200.utils import peaks  # noqa: E402

dev = torch.device("cuda:0")
out = {"gpu": torch.cuda.get_device_name(0)}
for dt, nm in ((torch.float64, "f64"), (torch.float32, "f32")):
    out[nm] = {"fma_peak_tflops": peaks.fma_peak_tflops(dev, dt), **peaks.special_weights(dev, dt)}
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "special_weights.json").write_text(json.dumps(out, indent=1))
print(json.dumps(out))
