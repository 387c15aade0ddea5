#!/usr/bin/env python3
"""Stand-alone M-step (stats_kernel + reduce + finalize) on the 10 M-trajectory synthetic set: CUDA-event time per call.
usage: stats_probe.py [n] [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import mtm_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda", 0)
eng = mtm_b200.Engine.get(dev)
z, x = bench.sample_global(n, 0, n, 8, 4, 4, bench.mixture_params(8, 4, 4), dev)
data = mtm_b200.DeviceData(z, x, dev, pattern_order=True)
lab = data.to_resident(bench.labels_global(n, 0, n, 8, dev))
eng.mstep(data, lab, 8, 0.0)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
torch.cuda.synchronize()
ev[0].record()
for k in range(reps):
    eng.mstep(data, lab, 8, 0.0)
    ev[k + 1].record()
torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[reps]) / reps
print(f"stand-alone M-step: n={n}  {ms:.3f} ms/call  {520 * n / ms * 1e-6:.1f} GB/s of algorithmic bytes")
