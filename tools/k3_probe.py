#!/usr/bin/env python3
"""Observation-only E-step (kernel 3: estep_kernel<4,4,HAS_Z=false>) on synthetic held-out trajectories, a few
launches with CUDA-event timing — the short command the ncu capture of that instantiation runs.
usage: k3_probe.py [n_trajectories] [launches]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import mtm_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda", 0)
eng = mtm_b200.Engine.get(dev)
comps = bench.mixture_params(8, 4, 4)
_, x = bench.sample_global(n, 0, n, 8, 4, 4, comps, dev, seed=987_000, want_z=False)
P = torch.from_numpy(np.stack([np.concatenate([[1.0 / 8], c["m"], c["S"].ravel(), c["A"].ravel(), c["G"].ravel(),
                                                c["H"].ravel(), c["L"].ravel()]) for c in comps])).to(dev)
data = mtm_b200.DeviceData(None, x, dev)
assign = torch.empty(n, dtype=torch.int32, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
eng.estep(data, P, 8, 8, assign=assign, d=4, want_info=False)
torch.cuda.synchronize()
ev[0].record()
for k in range(reps):
    eng.estep(data, P, 8, 8, assign=assign, d=4, want_info=False)
    ev[k + 1].record()
torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[reps]) / reps
print(f"obs-only E-step: n={n} n_c=8 T=8 l=4  {ms:.3f} ms/launch  {8 * n / ms * 1e-6:.3f} G evals/s")
