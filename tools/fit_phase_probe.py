"""Where the per-fit upload goes (bench.py e2e.per_fit.ms_upload_and_order): pinned -> HBM copy, pattern order,
permutes, each between device synchronisations, with and without a communicator attached.  Run alone or under
torchrun:  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/fit_phase_probe.py"""
import argparse
import ctypes
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trajectories", type=int, default=10_000_000)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    ctx = bench.Ctx(a)
    torch = ctx.torch
    T, d, l, n_c = 8, 4, 4, 8
    lo, hi = bench.shard_bounds(a.trajectories, ctx.world, ctx.rank)
    z, x = bench.sample_global(a.trajectories, lo, hi, T, d, l, bench.mixture_params(n_c, d, l), ctx.device)
    zh = torch.empty(z.shape, dtype=torch.float64, pin_memory=True).copy_(z)
    xh = torch.empty(x.shape, dtype=torch.float64, pin_memory=True).copy_(x)
    del z, x
    eng = ctx.eng

    def tick(label, t0, rows):
        torch.cuda.synchronize(ctx.device)
        t1 = time.perf_counter()
        rows.append((label, (t1 - t0) * 1e3))
        return t1

    for rep in range(a.reps):
        ctx.barrier()
        rows = []
        t = time.perf_counter()
        zd = zh.to(ctx.device)
        t = tick("h2d z (.to)", t, rows)
        xd = xh.to(ctx.device)
        t = tick("h2d x (.to)", t, rows)
        dd = ctx.mtm.DeviceData(zd, xd, ctx.device, pattern_order=False)
        t = tick("DeviceData(device tensors)", t, rows)
        eng._sync_stream()
        perm = torch.empty(dd.n, dtype=torch.int64, device=ctx.device)
        st = dd.struct()
        eng._chk(eng.lib.mtm_pattern_order(eng.h, ctypes.byref(st), ctypes.c_void_p(perm.data_ptr())))
        t = tick("mtm_pattern_order", t, rows)
        for name, src in (("z", zd), ("x", xd)):
            dst = torch.empty_like(src)
            t = tick(f"empty_like {name}", t, rows)
            eng._chk(eng.lib.mtm_permute_f64(eng.h, ctypes.c_void_p(src.data_ptr()), ctypes.c_void_p(dst.data_ptr()),
                                             T, dd.n, 4, ctypes.c_void_p(perm.data_ptr()), 0))
            t = tick(f"mtm_permute_f64 {name}", t, rows)
            del dst
        del dd, zd, xd, perm
        t = tick("free", t, rows)
        dd = ctx.mtm.DeviceData(zh, xh, ctx.device, pattern_order=True)
        t = tick("DeviceData(pinned, pattern_order) whole", t, rows)
        del dd
        print(f"rank {ctx.rank} rep {rep}: " + "; ".join(f"{k} {v:.1f} ms" for k, v in rows), file=sys.stderr, flush=True)
    ctx.barrier()


if __name__ == "__main__":
    main()
