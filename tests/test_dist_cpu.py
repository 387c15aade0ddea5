"""world_size-2 `gloo` tests (CPU) of the host-side sharding logic and of the property the
multi-GPU path relies on: sufficient statistics are additive over contiguous shards, so
all-reduce(statistics) followed by the same deterministic finalise on every rank gives every
rank the single-process parameters.  The arithmetic here is the oracle's (test infrastructure);
the CUDA twin of this test is tools/dist_parity.py (run on 2 GPUs, see profiles/)."""

import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def oracle_stats(z, x, assign, n_c):
    """Moments in the layout of mtm_stats_stride (csrc/common.cuh): per cluster
    [N | n0, sum z1, upper(sum z1 z1') | n1, sum u, upper(sum uu') | n2, sum w, upper(sum ww')]."""
    T, n, d = z.shape
    l = x.shape[2]

    def block(rows):
        k = rows.shape[1]
        iu = np.triu_indices(k)
        return np.concatenate([[rows.shape[0]], rows.sum(axis=0), (rows.T @ rows)[iu]])

    out = []
    for c in range(n_c):
        sel = assign == c
        Z, X = z[:, sel], x[:, sel]
        z0 = Z[0][np.isfinite(Z[0]).all(axis=1)]
        u = np.concatenate([Z[:-1].reshape(-1, d), Z[1:].reshape(-1, d)], axis=1)
        u = u[np.isfinite(u).all(axis=1)]
        w = np.concatenate([Z.reshape(-1, d), X.reshape(-1, l)], axis=1)
        w = w[np.isfinite(w).all(axis=1)]
        out.append(np.concatenate([[sel.sum()], block(z0), block(u), block(w)]))
    return np.stack(out)


def finalize_from_stats(S, d, l, n_total, alpha):
    """numpy twin of finalize_kernel (csrc/mstep.cu): parameters from moments."""
    rows = []
    for st in S:
        o = 1
        def take(k):
            nonlocal o
            cnt, s = st[o], st[o + 1:o + 1 + k]
            M = np.zeros((k, k))
            M[np.triu_indices(k)] = st[o + 1 + k:o + 1 + k + k * (k + 1) // 2]
            M = M + np.triu(M, 1).T
            o += 1 + k + k * (k + 1) // 2
            return cnt, s, M
        n0, s0, M0 = take(d)
        m = s0 / n0
        Sg = (M0 - n0 * np.outer(m, m)) / (n0 - 1)
        def reg(cnt, s, M, p, q):
            B = np.linalg.solve(M[:p, :p] + alpha * np.eye(p), M[:p, p:])
            rbar = (s[p:] - s[:p] @ B) / cnt
            ssr = M[p:, p:] - B.T @ M[:p, p:] - M[:p, p:].T @ B + B.T @ M[:p, :p] @ B
            return B, (ssr - cnt * np.outer(rbar, rbar)) / (cnt - 1)
        A, G = reg(*take(2 * d), d, d)
        H, L = reg(*take(d + l), d, l)
        rows.append(np.concatenate([[st[0] / n_total], m, Sg.ravel(), A.ravel(), G.ravel(), H.ravel(), L.ravel()]))
    return np.stack(rows)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mtm_b200
    from mtm_b200 import dist as mdist
    from oracle import dense as od

    g = np.load(os.path.join(ROOT, "tests", "golden", "em_traces.npz"))
    key = "c2s"
    n, T, d, l, n_c, seed, _ = (int(v) for v in g[f"{key}_meta"])
    z, x = g[f"{key}_z"], g[f"{key}_x"]
    lo, hi = mdist.shard_bounds(n, world, rank)
    zs, xs = mdist.shard_arrays(z, x, world, rank)
    assert zs.shape[1] == hi - lo and zs.flags.c_contiguous
    a_loc = mdist.global_random_init(n, n_c, seed, world, rank)
    assert np.array_equal(a_loc, g[f"{key}_assign0"][lo:hi])
    # local statistics -> all-reduce -> identical finalise on every rank
    st = torch.from_numpy(oracle_stats(zs, xs, a_loc, n_c))
    tail = torch.tensor([0.0, float(hi - lo), 0.0, 0.0], dtype=torch.float64)  # [switches | n | NONFINITE | SINGULAR]
    buf = torch.cat([st.reshape(-1), tail])
    dist.all_reduce(buf)
    stride = mtm_b200._capi.load().mtm_stats_stride(d, l)
    assert st.shape[1] == stride
    n_total = int(buf[-3].item())
    P = finalize_from_stats(buf[:-4].reshape(n_c, stride).numpy(), d, l, n_total, 0.0)
    ref = od.m_step(z, x, g[f"{key}_assign0"], n_c, 0.0).packed()
    err = float(np.max(np.abs(P - ref) / np.maximum(1e-6, np.abs(ref))))
    # gather of ragged label shards
    full = mdist.gather_assignments(torch.from_numpy(a_loc.astype(np.int32)), n, world, rank)
    ok = err < 1e-9 and n_total == n and np.array_equal(full, g[f"{key}_assign0"])
    # host-side helpers of the sharded drop-in class: ragged all-gather of the last axis, integer sums, group probe
    part = torch.arange(lo, hi, dtype=torch.float64).repeat(3, 1) * torch.tensor([[1.0], [2.0], [3.0]])
    whole = mdist.gather_last(part, n, world, rank)
    ok = ok and tuple(whole.shape) == (3, n) and bool(torch.equal(whole[1], 2.0 * torch.arange(n, dtype=torch.float64)))
    ok = ok and mdist.sum_int(rank + 1, "cpu") == world * (world + 1) // 2
    ok = ok and mdist.active_group() == (world, rank) and mdist.active_group(False) == (1, 0)
    # every rank ends with bit-identical parameters
    mine = torch.from_numpy(P.copy())
    other = mine.clone()
    dist.broadcast(other, 0)
    ok = ok and bool(torch.equal(mine, other))
    q.put((rank, ok, err))
    dist.destroy_process_group()


def test_shard_bounds_partition():
    sys.path.insert(0, ROOT)
    from mtm_b200 import dist as mdist

    for n in (1, 7, 128, 1000, 10_000_001):
        for world in (1, 2, 3, 8):
            b = [mdist.shard_bounds(n, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_stats_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
