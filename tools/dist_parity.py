#!/usr/bin/env python3
"""Multi-GPU parity check, run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tools/dist_parity.py

Every rank takes a contiguous shard of a golden data set, the ranks fit together (one all-reduce of
the sufficient statistics per EM iteration inside mtm_em_fit) and the result must equal the
reference's single-process trace: switch counts and assignments exactly, parameters to 1e-9."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mtm_b200  # noqa: E402
from mtm_b200 import dist as mdist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    eng = mtm_b200.Engine.get(device)
    mdist.attach(eng)
    g = np.load(os.path.join(ROOT, "tests", "golden", "em_traces.npz"))
    ok = True
    for key in ("c2s", "c3s", "slow", "adnish", "short", "guardexit"):
        n, T, d, l, n_c, seed, n_steps = (int(v) for v in g[f"{key}_meta"])
        z, x, alpha = g[f"{key}_z"], g[f"{key}_x"], float(g[f"{key}_alpha"])
        lo, hi = mdist.shard_bounds(n, world, rank)
        # every rank keeps its shard in pattern order (the product path); labels go back to the caller's order
        data = mtm_b200.DeviceData(z[:, lo:hi], x[:, lo:hi], device, pattern_order=True)
        assign = data.to_resident(g[f"{key}_assign0"][lo:hi].astype(np.int32))
        params, info = eng.em_fit(data, assign, n_c, alpha, n_steps)
        full = mdist.gather_assignments(data.from_resident(assign), n, world, rank)
        P = params.cpu().numpy()
        err = float(np.max(np.abs(P - g[f"{key}_params"]) / np.maximum(1e-6, np.abs(g[f"{key}_params"]))))
        good = (info["switches"] == list(g[f"{key}_switches"]) and np.array_equal(full, g[f"{key}_assign"])
                and err < 1e-9)
        # scores are all-reduced too
        if key != "guardexit":
            q, ll, _ = eng.score(data, params, n_c)
            good = good and abs(q - float(g[f"{key}_Q"])) < 1e-9 * abs(float(g[f"{key}_Q"]))
            good = good and abs(ll - float(g[f"{key}_LL"])) < 1e-9 * abs(float(g[f"{key}_LL"]))
        # every rank must hold bit-identical parameters
        ref = params.clone()
        dist.broadcast(ref, 0)
        same = bool(torch.equal(torch.nan_to_num(ref), torch.nan_to_num(params)))
        ok = ok and good and same
        if rank == 0:
            print(f"{key:10s} world={world} exit={info['exit']:10s} E-steps={info['n_e_steps']:3d} "
                  f"max rel param err {err:.1e} identical-across-ranks={same} -> {'OK' if good else 'MISMATCH'}")
    # batched restarts (mtm_sweep_fit) on shards: Q of every restart must equal the reference's objectives
    gr = np.load(os.path.join(ROOT, "tests", "golden", "restarts.npz"))
    z, x = gr["B_z"], gr["B_x"]
    n, T, d, l, n_c, n_starts = (int(v) for v in gr["B_meta"])
    lo, hi = mdist.shard_bounds(n, world, rank)
    data = mtm_b200.DeviceData(z[:, lo:hi], x[:, lo:hi], device, pattern_order=True)
    a0 = np.stack([np.random.default_rng(100 + i).integers(0, n_c, n)[lo:hi] for i in range(n_starts)]).astype(np.int32)
    assign = data.to_resident(a0)
    P, infos, scores, sstat = eng.sweep_fit(data, assign, n_c, 0.0, 100)
    ref_obj = gr["B_objectives"][1:]  # [0] is the k-means start
    err = float(np.max(np.abs(scores[:, 0] - ref_obj) / np.abs(ref_obj)))
    good = err < 1e-9 and all(i["exit"] in ("converged", "guard", "exhausted") for i in infos)
    ok = ok and good
    if rank == 0:
        print(f"sweep      world={world} restarts={n_starts} max rel Q err {err:.1e} -> {'OK' if good else 'MISMATCH'}")
    flag = torch.tensor([1 if ok else 0], device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    mdist.detach(eng)
    dist.destroy_process_group()
    if rank == 0:
        print("DIST PARITY", "PASS" if int(flag.item()) else "FAIL")
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
