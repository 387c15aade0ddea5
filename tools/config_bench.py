#!/usr/bin/env python3
"""Timings of the other BASELINE.json configurations on ONE B200 (they are parity-test cases, not the
bench line; the numbers go into DESIGN.md / profiles/).

  C1  ADNI-shaped (n=513, T=4, d=2, l=4, n_c=4, alpha=1): train_with_multiple_random_starts(n_starts=1000)
  C2  1 M trajectories, n_c=4: EM iteration
  C4  observation-only predict (hidden states marginalised), n_c=8, on --c4-n trajectories
  C5  restart sweep n_c=2..12 x 32 restarts on --c5-n trajectories (batched, kernel 4)
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import mtm_b200  # noqa: E402
from mtm_b200.framework import marginalizable_mixture_model as mixmodel  # noqa: E402


def timed(fn, reps=3):
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--c4-n", type=int, default=100_000_000)
    ap.add_argument("--c5-n", type=int, default=1_000_000)
    ap.add_argument("--c1-starts", type=int, default=1000)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    eng = mtm_b200.Engine.get(dev)
    out = {}

    # ---- C1 ----
    g = np.load(os.path.join(ROOT, "tests", "golden", "adni_fold0.npz"))
    ztr, xtr = g["ztr"], g["xtr"]
    proto = mixmodel.MMLinGaussSS_marginalizable(n_clusters=4, states=ztr, observations=xtr, init="k-means", alpha=1.0)
    proto.train_with_multiple_random_starts(n_starts=8, use_cache=False)  # warm-up
    t0 = time.perf_counter()
    best = proto.train_with_multiple_random_starts(n_starts=args.c1_starts, use_cache=False)
    dt = time.perf_counter() - t0
    out["C1"] = dict(n=int(ztr.shape[1]), n_starts=args.c1_starts, seconds=dt, fits_per_s=(args.c1_starts + 1) / dt,
                     best_seed=int(best.random_seed), best_Q=float(best.e_complete_data_log_lik()),
                     note="reference: ~0.68 s per fit on one core (SURVEY 6) -> ~680 s for 1000 starts")

    # ---- C2 ----
    comps = bench.mixture_params(4, 4, 4)
    z, x = bench.sample_device(1_000_000, 8, 4, 4, comps, seed=4321, device=dev)
    data = mtm_b200.DeviceData(z, x, dev, pattern_order=True)
    a = torch.randint(0, 4, (1_000_000,), device=dev, dtype=torch.int32)
    p, info = eng.em_fit(data, a.clone(), 4, 0.0, 3)
    t = timed(lambda: eng.em_fit(data, a.clone(), 4, 0.0, 10), reps=3)
    out["C2"] = dict(n=1_000_000, n_c=4, ms_per_em_iteration=1e3 * t / 10, evals_per_s=4e6 * 10 / t)

    # ---- C5 (reduced n) ----
    n5 = args.c5_n
    comps8 = bench.mixture_params(8, 4, 4)
    z5, x5 = bench.sample_device(n5, 8, 4, 4, comps8, seed=99, device=dev)
    d5 = mtm_b200.DeviceData(z5, x5, dev, pattern_order=True)
    tot_sets, t5 = 0, 0.0
    steps = 3
    for n_c in range(2, 13):
        a5 = torch.randint(0, n_c, (32, n5), device=dev, dtype=torch.int32)
        eng.sweep_fit(d5, a5.clone(), n_c, 0.0, 1, want_scores=False)  # warm-up: scratch buffers of this n_c
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.sweep_fit(d5, a5, n_c, 0.0, steps, want_scores=False)
        torch.cuda.synchronize()
        t5 += time.perf_counter() - t0
        tot_sets += 32 * n_c
    out["C5"] = dict(n=n5, restarts=32, n_c="2..12", parameter_sets=tot_sets, em_iterations=steps,
                     seconds_per_sweep_iteration=t5 / steps, evals_per_s=tot_sets * n5 * steps / t5)
    del z5, x5, d5, z, x, data
    torch.cuda.empty_cache()

    # ---- C4 ----
    n4 = args.c4_n
    g4 = torch.Generator(device=dev)
    g4.manual_seed(7)
    x4 = torch.randn((8, n4, 4), generator=g4, device=dev, dtype=torch.float64)
    chunk = 10_000_000
    for s in range(0, n4, chunk):  # 20 % missing, chunked to bound temporaries
        m = torch.rand((8, min(chunk, n4 - s), 4), generator=g4, device=dev) < 0.2
        x4[:, s:s + m.shape[1]][m] = float("nan")
        del m
    d4 = mtm_b200.DeviceData(None, x4, dev)
    P = torch.from_numpy(np.stack([np.concatenate([[1.0 / 8], c["m"], c["S"].ravel(), c["A"].ravel(), c["G"].ravel(),
                                                    c["H"].ravel(), c["L"].ravel()]) for c in comps8])).to(dev)
    assign = torch.empty(n4, dtype=torch.int32, device=dev)
    eng.estep(d4, P, 8, 8, assign=assign, d=4, want_info=False)
    t4 = timed(lambda: eng.estep(d4, P, 8, 8, assign=assign, d=4, want_info=False), reps=3)
    out["C4"] = dict(n=n4, n_c=8, ms=1e3 * t4, evals_per_s=8 * n4 / t4, hbm_gbs=(256 + 4) * n4 / t4 * 1e-9)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
