"""Development helper: the two support-tagged warp-slot builds of a DiffTRe chunk, frame-resident route against the
multi-launch route (MYTHOS_B200_NL_FRAME=0), timed with CUDA events."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from mythos_b200.utils import neighbors, synthetic

F = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
dev = torch.device("cuda:0")
s = synthetic.assembly(17, seed=1)
c, q = synthetic.rejittered_frames(s, F)
c32 = torch.tensor(c, device=dev, dtype=torch.float32)
bonded = torch.tensor(s.topology.bonded_neighbors)
n = c32.shape[1]
wpf = (n + 31) // 32
for r_cut, lane_slots, W in ((float(sys.argv[2]) if len(sys.argv) > 2 else 1.68, 28, 320), (float(sys.argv[3]) if len(sys.argv) > 3 else 2.36, 52, 680)):
    cap = wpf * W
    pairs = torch.empty((F, 2, cap), dtype=torch.int32, device=dev)
    count = torch.empty((F,), dtype=torch.int32, device=dev)
    ov = torch.zeros((1,), dtype=torch.int32, device=dev)
    mr = torch.empty((F, 2), dtype=torch.int32, device=dev)
    ws = None
    for route, blk in (("0", ""), ("1", "256"), ("1", "384"), ("1", "512"), ("1", "")):
        os.environ["MYTHOS_B200_NL_FRAME"] = route
        os.environ["MYTHOS_B200_NL_FRAME_BLOCK"] = blk
        fn = lambda: neighbors.build_pairs(c32, bonded, (0.0, 0.0, 0.0), r_cut, 0.0, cap, ws, tag_bits=1 << 30, out=(pairs, count, ov), max_row=mr, warp_slots=(lane_slots, 0, W))
        ws = fn()[3]; fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(f"cutoff {r_cut}: route {route} block {blk or 'auto'}: {np.median(ts):.3f} ms per {F} frames; pairs/frame {float(count.float().mean()):.0f} "
              f"max lane/warp {mr.max(0).values.tolist()} overflow {int(ov.item())}")
