#!/usr/bin/env python3
"""Probe: batched restarts (mtm_sweep_fit) on the caller's order vs the pattern-ordered resident set."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import mtm_b200  # noqa: E402

dev = torch.device("cuda", 0)
eng = mtm_b200.Engine.get(dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
comps = bench.mixture_params(8, 4, 4)
z, x = bench.sample_device(n, 8, 4, 4, comps, seed=99, device=dev)
for order in (False, True, False, True):
    d = mtm_b200.DeviceData(z, x, dev, pattern_order=order)
    for n_c in (4, 8, 12):
        a = torch.randint(0, n_c, (32, n), device=dev, dtype=torch.int32)
        a = d.to_resident(a)
        eng.sweep_fit(d, a.clone(), n_c, 0.0, 1, want_scores=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.sweep_fit(d, a, n_c, 0.0, 3, want_scores=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"pattern_order={order} n_c={n_c:2d}: {1e3 * dt / 3:8.2f} ms per sweep iteration (32 restarts)", flush=True)
