#!/usr/bin/env python3
"""BASELINE.json configs 4 and 5 at their named sizes, sharded over the GPUs of one box (run under torchrun):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
        tools/config_bench_dist.py

C4: measurement-only predict (hidden states marginalised) of 100 M held-out trajectories: no collective.
C5: model-selection sweep n_c = 2..12 x 32 random restarts on 10 M trajectories: batched restarts
    (mtm_sweep_fit), one all-reduce of the widened statistics per sweep iteration.
Times are device-synchronised wall clock between barriers, max over ranks; one JSON object on rank 0."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import mtm_b200  # noqa: E402
from mtm_b200 import dist as mdist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--c4-n", type=int, default=100_000_000)
    ap.add_argument("--c5-n", type=int, default=10_000_000)
    ap.add_argument("--c5-steps", type=int, default=2)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group("nccl", device_id=dev)
    eng = mtm_b200.Engine.get(dev)
    mdist.attach(eng)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn):
        barrier()
        t0 = time.perf_counter()
        fn()
        barrier()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    out = {"world": world}
    comps8 = bench.mixture_params(8, 4, 4)

    # ---- C5 ----
    n5 = args.c5_n // world
    z5, x5 = bench.sample_device(n5, 8, 4, 4, comps8, seed=99 + rank, device=dev)
    d5 = mtm_b200.DeviceData(z5, x5, dev, pattern_order=True)
    del z5, x5
    tot_sets, t5 = 0, 0.0
    for n_c in range(2, 13):
        a5 = torch.randint(0, n_c, (32, n5), device=dev, dtype=torch.int32)
        eng.sweep_fit(d5, a5.clone(), n_c, 0.0, 1, want_scores=False)  # warm-up: scratch buffers of this n_c
        t5 += timed(lambda: eng.sweep_fit(d5, a5, n_c, 0.0, args.c5_steps, want_scores=False))
        tot_sets += 32 * n_c
        del a5
    out["C5"] = dict(n_total=n5 * world, restarts=32, n_c="2..12", parameter_sets=tot_sets, em_iterations=args.c5_steps,
                     seconds_per_sweep_iteration=t5 / args.c5_steps,
                     evals_per_s=tot_sets * n5 * world * args.c5_steps / t5,
                     note="each sweep_fit call also runs the first M-step of its 32 restarts")
    del d5
    torch.cuda.empty_cache()

    # ---- C4 ----
    n4 = args.c4_n // world
    g4 = torch.Generator(device=dev)
    g4.manual_seed(7 + rank)
    x4 = torch.randn((8, n4, 4), generator=g4, device=dev, dtype=torch.float64)
    chunk = 5_000_000
    for s in range(0, n4, chunk):
        m = torch.rand((8, min(chunk, n4 - s), 4), generator=g4, device=dev) < 0.2
        x4[:, s:s + m.shape[1]][m] = float("nan")
        del m
    d4 = mtm_b200.DeviceData(None, x4, dev)
    P = torch.from_numpy(np.stack([np.concatenate([[1.0 / 8], c["m"], c["S"].ravel(), c["A"].ravel(), c["G"].ravel(),
                                                    c["H"].ravel(), c["L"].ravel()]) for c in comps8])).to(dev)
    assign = torch.empty(n4, dtype=torch.int32, device=dev)
    eng.estep(d4, P, 8, 8, assign=assign, d=4, want_info=False)
    reps = 3

    def run4():
        for _ in range(reps):
            eng.estep(d4, P, 8, 8, assign=assign, d=4, want_info=False)

    t4 = timed(run4) / reps
    out["C4"] = dict(n_total=n4 * world, n_c=8, ms=1e3 * t4, evals_per_s=8 * n4 * world / t4,
                     hbm_gbs_per_gpu=(256 + 4) * n4 / t4 * 1e-9)
    if rank == 0:
        print(json.dumps(out, indent=1))
    mdist.detach(eng)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
