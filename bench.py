#!/usr/bin/env python3
"""Benchmark of the hard-assignment EM hot path (BASELINE.json: trajectory x cluster log-lik
evaluations / s and EM iterations / s at 10 M trajectories).

    python bench.py --gpus N --steps K --warmup W            # this implementation (B200)
    python bench.py --impl reference --steps K --warmup W     # CPU port of the reference algorithm

A *step* is one EM iteration over the whole data set: E-step (n x n_c log-likelihood
evaluations + argmax + switch count) followed by the M-step (sufficient statistics, the
all-reduce when N > 1, parameter solve).  Exactly K steps are timed, with no convergence exit.
Workload at N = 1: config C3 of BASELINE.json (10 M trajectories, T = 8, d = l = 4, n_c = 8,
20 % element-wise missing); N > 1 shards THE SAME 10 M trajectories (strong scaling): the data set
is generated in 64 seeded chunks of the global trajectory axis, a rank generates the chunks of its
contiguous shard, so every N sees bit-identical data and initial labels.

The timed workload is stationary: after the W warm-up iterations (real EM progress from the random
initial assignment) the state (labels, parameters) is saved and restored before every timed step
(two device-to-device copies, inside the timed region), so every timed step is the same EM
iteration W+1 — the measurement cannot drift into a collapsed or converged model.  The run is
refused (exit code 1) if that iteration leaves a cluster with <= 3 members (where the reference's
guard, MMM:1346-1356, would have stopped) or raises a numerical status.

Besides the headline the default run measures the other BASELINE configurations as `configs`:
C1 (the ADNI fold with 1000 restarts through the class, against the reference's golden fold),
C2 (1 M x n_c 4, EM iteration), C4 (observation-only predict of 100 M held-out trajectories) and
C5 (restart sweep n_c = 2..12 x 32 restarts on the 10 M set), each with its own clocks, roofline
figure and oracle parity gate.

One JSON line on stdout (rank 0).
"""

from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "traj_x_cluster_loglik_evals_per_s"
UNIT = "evals/s"
G_CHUNKS = 64  # chunks of the global trajectory axis, each with its own seed


# -------------------------------------------------------------------------------------------------
# synthetic mixture of linear-Gaussian state-space models
# -------------------------------------------------------------------------------------------------
def mixture_params(n_c, d, l, seed=1234, sep=0.3):
    """n_c clusters = a common base model moved `sep` of the way towards independent random
    models (SURVEY 8d parameter styles).  Fully independent clusters are so well separated that
    hard EM empties a cluster within 2-3 iterations (then the reference itself raises)."""
    rng = np.random.default_rng(seed)

    def spd(k):
        B = rng.normal(size=(k, k))
        return B @ B.T / k + 0.3 * np.eye(k)

    def comp():
        return dict(m=2 * rng.normal(size=d), S=spd(d), A=rng.normal(scale=0.5, size=(d, d)), G=spd(d),
                    H=rng.normal(size=(d, l)), L=spd(l))

    base = comp()
    out = []
    for _ in range(n_c):
        p = comp()
        out.append({k: base[k] + sep * (p[k] - base[k]) for k in base})
    return out


def sample_host(n, T, d, l, comps, seed, p_missing=0.2):
    """numpy sampler (CPU arms and small parity prefixes)."""
    rng = np.random.default_rng(seed)
    lab = rng.integers(0, len(comps), size=n)
    z, x = np.empty((T, n, d)), np.empty((T, n, l))
    for c, p in enumerate(comps):
        idx = np.flatnonzero(lab == c)
        cs, cg, cl = (np.linalg.cholesky(p[k]) for k in ("S", "G", "L"))
        zc = np.empty((T, idx.size, d))
        zc[0] = p["m"] + rng.normal(size=(idx.size, d)) @ cs.T
        for t in range(T):
            if t:
                zc[t] = zc[t - 1] @ p["A"] + rng.normal(size=(idx.size, d)) @ cg.T
            x[t, idx] = zc[t] @ p["H"] + rng.normal(size=(idx.size, l)) @ cl.T
        z[:, idx] = zc
    z[rng.random(z.shape) < p_missing] = np.nan
    x[rng.random(x.shape) < p_missing] = np.nan
    return z, x


def sample_device(n, T, d, l, comps, seed, device, p_missing=0.2):
    """Same generative model on the device (torch, seeded): 10 M trajectories in ~1 s."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    f64 = dict(dtype=torch.float64, device=device)
    lab = torch.randint(0, len(comps), (n,), generator=g, device=device)
    z = torch.empty((T, n, d), **f64)
    x = torch.empty((T, n, l), **f64)
    tt = lambda a: torch.as_tensor(a, **f64)
    for c, p in enumerate(comps):
        idx = torch.nonzero(lab == c).squeeze(1)
        k = idx.numel()
        cs, cg, cl = (tt(np.linalg.cholesky(p[key])) for key in ("S", "G", "L"))
        A, H, m = tt(p["A"]), tt(p["H"]), tt(p["m"])
        zc = m + torch.randn((k, d), generator=g, **f64) @ cs.T
        for t in range(T):
            if t:
                zc = zc @ A + torch.randn((k, d), generator=g, **f64) @ cg.T
            z[t, idx] = zc
            x[t, idx] = zc @ H + torch.randn((k, l), generator=g, **f64) @ cl.T
    z[torch.rand(z.shape, generator=g, device=device) < p_missing] = float("nan")
    x[torch.rand(x.shape, generator=g, device=device) < p_missing] = float("nan")
    return z, x


def chunk_bounds(n_total, g):
    return g * n_total // G_CHUNKS, (g + 1) * n_total // G_CHUNKS


def shard_bounds(n_total, world, rank):
    return rank * n_total // world, (rank + 1) * n_total // world


def sample_global(n_total, lo, hi, T, d, l, comps, device, seed=4321, want_z=True):
    """Rows [lo, hi) of the GLOBAL synthetic data set of n_total trajectories: chunk g of the trajectory axis is
    sample_device(seed + g).  The same rows come out whatever the shard boundaries are, so every GPU count
    works on the same data."""
    import torch

    f64 = dict(dtype=torch.float64, device=device)
    z = torch.empty((T, hi - lo, d), **f64) if want_z else None
    x = torch.empty((T, hi - lo, l), **f64)
    for g in range(G_CHUNKS):
        c0, c1 = chunk_bounds(n_total, g)
        a, b = max(lo, c0), min(hi, c1)
        if a >= b:
            continue
        zc, xc = sample_device(c1 - c0, T, d, l, comps, seed + g, device)
        x[:, a - lo:b - lo] = xc[:, a - c0:b - c0]
        if want_z:
            z[:, a - lo:b - lo] = zc[:, a - c0:b - c0]
        del zc, xc
    return z, x


def labels_global(n_total, lo, hi, n_c, device, seed=42, rows=None):
    """Rows [lo, hi) of the global initial assignment (int32, caller's order); rows = R gives (R, hi - lo)."""
    import torch

    out = torch.empty((hi - lo,) if rows is None else (rows, hi - lo), dtype=torch.int32, device=device)
    for g in range(G_CHUNKS):
        c0, c1 = chunk_bounds(n_total, g)
        a, b = max(lo, c0), min(hi, c1)
        if a >= b:
            continue
        gen = torch.Generator(device=device)
        gen.manual_seed(seed + g)
        lab = torch.randint(0, n_c, (c1 - c0,) if rows is None else (rows, c1 - c0), generator=gen, device=device,
                            dtype=torch.int32)
        out[..., a - lo:b - lo] = lab[..., a - c0:b - c0]
    return out


# -------------------------------------------------------------------------------------------------
# algorithmic work (DESIGN.md "Roofline accounting")
# -------------------------------------------------------------------------------------------------
def flops_per_eval(T, d, l, kz_mean, kx_mean):
    """fp64 flops (FMA = 2) of ONE (trajectory, cluster) evaluation by the sequential algorithm,
    counting only arithmetic on PRESENT scalars (work the kernel spends on predicated-off missing
    scalars and the Newton steps of the reciprocal are not counted).
      per step      x-predictive  2dl + 2d^2 l + 2 d l(l+1)/2
      per step > 1  propagate     2d^2 + 2d^3 + 2 d d(d+1)/2
      per present z scalar        6 + d + 2d + 2 d(d+1)/2
      per present x scalar        6 + d + 2d + 2 d(d+1)/2 + (l-1)/2 (2d + 3) + (l^2-1)/3
    """
    pred = 2 * d * l + 2 * d * d * l + d * l * (l + 1)
    prop = 2 * d * d + 2 * d ** 3 + d * d * (d + 1)
    per_z = 6 + 3 * d + d * (d + 1)
    per_x = 6 + 3 * d + d * (d + 1) + 0.5 * (l - 1) * (2 * d + 3) + (l * l - 1) / 3.0
    return T * pred + (T - 1) * prop + kz_mean * per_z + kx_mean * per_x


def flops_per_eval_survey(T, d, l, k_mean):
    """250 T + 250 (T-1) + 101 k at d = l = 4 (full (d+l)-dim rank-1 update per
    present scalar): the count of a less structured sequential algorithm, quoted for comparison."""
    D = d + l
    pred = 2 * d * l + 2 * d * d * l + (2 * d + 1) * l * (l + 1) // 2
    prop = 2 * d * d + 2 * d ** 3 + (2 * d + 1) * d * (d + 1) // 2
    return T * pred + (T - 1) * prop + k_mean * (D * (D + 1) + 3 * D + 5)


def bytes_per_traj(T, d, l):
    """HBM bytes one EM iteration must move per trajectory: the data once + old/new label."""
    return 8 * T * (d + l) + 8


# -------------------------------------------------------------------------------------------------
# clocks during the timed region
# -------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.index, self.samples, self.reasons, self._stop, self._thr = index, [], set(), threading.Event(), None
        self.max_mhz = None

    def _run(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20,
                     "hw_thermal_slowdown": 0x40, "hw_power_brake": 0x80, "sync_boost": 0x10,
                     "applications_clocks": 0x2, "display_clocks": 0x100}
            while not self._stop.is_set():
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                try:
                    r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:  # noqa: BLE001
                    r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
                time.sleep(0.02)
        except Exception as e:  # noqa: BLE001
            self.reasons.add(f"sampler_error:{type(e).__name__}")

    def __enter__(self):
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thr.join(timeout=2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "samples": len(s),
                "reasons": sorted(self.reasons)}


# -------------------------------------------------------------------------------------------------
# CPU arm: port of the reference algorithm (oracle/), bounded sample
# -------------------------------------------------------------------------------------------------
def host_threads():
    """Cores this process may use.  Passed EXPLICITLY to the C port: torchrun exports OMP_NUM_THREADS=1, which
    must not starve the CPU arm at N > 1."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_em_iteration(z, x, assign, n_c, par, nthreads):
    """One EM iteration the way the reference does it: dense joint covariance per cluster,
    per-row det + solve by LU (C port, OpenMP over rows), numpy M-step (oracle/dense.py)."""
    from oracle import cport as oc
    from oracle import dense as od

    ll = oc.dense_loglik_matrix(par, z, x, nthreads=nthreads)
    new = od.assign_from_loglik(par.pi, ll)
    return new, od.m_step(z, x, new, n_c, 0.0)


def run_cpu_arm(args, cfg, steps, warmup):
    from oracle import dense as od

    n_s = args.cpu_sample
    comps = mixture_params(cfg["n_c"], cfg["d"], cfg["l"])
    z, x = sample_host(n_s, cfg["T"], cfg["d"], cfg["l"], comps, seed=4321)
    assign = od.random_init(n_s, cfg["n_c"], 42)
    par = od.m_step(z, x, assign, cfg["n_c"], 0.0)
    threads = host_threads()
    for _ in range(warmup):
        assign, par = cpu_em_iteration(z, x, assign, cfg["n_c"], par, threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        assign, par = cpu_em_iteration(z, x, assign, cfg["n_c"], par, threads)
    dt = time.perf_counter() - t0
    value = n_s * cfg["n_c"] * steps / dt
    return dict(value=value, unit=UNIT, cores=threads, kind="port",
                sample=f"{n_s} trajectories of the same synthetic workload x {steps} EM iterations "
                       f"(dense-covariance LU port of SSM:728-773 in C with OpenMP, {threads} threads set "
                       f"explicitly + numpy M-step)",
                ms_per_step=1e3 * dt / steps, host_cpus=os.cpu_count())


# -------------------------------------------------------------------------------------------------
# helpers of the B200 arm
# -------------------------------------------------------------------------------------------------
def rel_err(a, b, floor=1e-6):
    """max |a - b| / max(|b|, floor): relative error with an explicit absolute floor for entries near zero."""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    with np.errstate(all="ignore"):
        d = np.abs(a - b) / np.maximum(np.abs(b), floor)
    d[np.isnan(a) & np.isnan(b)] = 0.0
    d[(a == b)] = 0.0
    return float(np.max(d)) if d.size else 0.0


class Ctx:
    """Everything a leg needs: engine, ranks, device, timing helpers."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        import mtm_b200

        self.torch, self.dist, self.mtm = torch, dist, mtm_b200
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.device = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.device)
        self.eng = mtm_b200.Engine.get(self.device)
        if self.world > 1:
            from mtm_b200 import dist as mdist

            mdist.attach(self.eng)  # library-owned NCCL communicator + peer-memory mailboxes
        self.fp64_peak = None
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            pass
        self.hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        self.hbm_src = "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.device)

    def max_over_ranks(self, vals):
        if self.world == 1:
            return [float(v) for v in vals]
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t.cpu()]

    def sum_over_ranks(self, t):
        if self.world > 1:
            self.dist.all_reduce(t)
        return t

    def event(self):
        return self.torch.cuda.Event(enable_timing=True)

    def gather_labels(self, local_caller_order, n_total):
        """(n_total,) int32 numpy on every rank from the per-rank shards (caller's order)."""
        torch = self.torch
        if self.world == 1:
            return local_caller_order.cpu().numpy()
        sizes = [shard_bounds(n_total, self.world, r)[1] - shard_bounds(n_total, self.world, r)[0]
                 for r in range(self.world)]
        pad = max(sizes)
        buf = torch.full((pad,), -1, dtype=torch.int32, device=self.device)
        buf[: local_caller_order.numel()] = local_caller_order
        out = [torch.empty_like(buf) for _ in range(self.world)]
        self.dist.all_gather(out, buf)
        return np.concatenate([o[:s].cpu().numpy() for o, s in zip(out, sizes)])


def data_stats(torch, z, x):
    """present scalars per trajectory and the point-step statistics the flop accounting needs (local shard)."""
    n = x.shape[1]
    kx = float(torch.isfinite(x).sum().item()) / n
    if z is None:
        return dict(kz=0.0, kx=kx)
    kz = float(torch.isfinite(z).sum().item()) / n
    out = dict(kz=kz, kx=kx)
    return out


def warp_step_classes(torch, z_resident):
    """Fractions of the (warp, time step) pairs of the resident order that take each code path of the E-step:
    point mass (no hidden-state coordinate missing in any of the 32 trajectories), rank one (at most one per
    trajectory), generic."""
    T, n, _ = z_resident.shape
    miss = (~torch.isfinite(z_resident)).sum(dim=2)                      # (T, n)
    pad = (-n) % 32
    if pad:
        miss = torch.cat([miss, miss[:, -1:].expand(T, pad)], dim=1)
    worst = miss.reshape(T, -1, 32).amax(dim=2)                          # (T, warps)
    tot = float(worst.numel())
    return {"point": float((worst == 0).sum().item()) / tot, "rank_one": float((worst == 1).sum().item()) / tot,
            "generic": float((worst >= 2).sum().item()) / tot}


def oracle_gate(ctx, z, x, params, n_c, T, npre=4096, d=None):
    """E-step of a caller-order prefix against the sequential C oracle (ll, assignments, closest call)."""
    from oracle import cport as oc

    sub = ctx.mtm.DeviceData(None if z is None else z[:, :npre].contiguous(), x[:, :npre].contiguous(), ctx.device)
    r = ctx.eng.estep(sub, params, n_c, T, want_ll=True, d=d)
    P = params.cpu().numpy()
    llo = oc.seq_loglik_matrix(P, None if z is None else sub.z.cpu().numpy(), sub.x.cpu().numpy())
    ll = r["ll"].cpu().numpy()
    with np.errstate(all="ignore"):
        pren = np.sort(np.log(P[:, :1]) + llo, axis=0)  # log pi_c + ll_c, ascending over clusters
        margin = float(np.nanmin(pren[-1] - pren[-2])) if n_c > 1 else None
    return {"trajectories": int(sub.n), "max_rel_ll_err": rel_err(ll, llo),
            "assignments_equal": bool(np.array_equal(r["assign"].cpu().numpy(), oc.assign(llo, P[:, 0]))),
            # smallest gap between the best and the second-best cluster of a trajectory (log domain): how
            # far the closest call in the sample is from a tie that rounding could flip
            "min_top2_margin": margin, "status": int(r["status"])}


class EmRunner:
    """A resident shard + labels + parameters, stepped through the C ABI (mtm_mstep / mtm_em_step)."""

    def __init__(self, ctx, z, x, assign_caller, n_c, pattern_order=True):
        import ctypes

        torch = ctx.torch
        self.ctx, self.n_c, self.ct = ctx, n_c, ctypes
        t0, t1 = ctx.event(), ctx.event()
        t0.record()
        self.data = ctx.mtm.DeviceData(z, x, ctx.device, pattern_order=pattern_order)
        t1.record()
        torch.cuda.synchronize(ctx.device)
        self.order_ms = t0.elapsed_time(t1)
        self.assign = self.data.to_resident(assign_caller)
        d, l = self.data.d, self.data.l
        self.params = torch.empty((n_c, ctx.eng.param_stride(d, l)), dtype=torch.float64, device=ctx.device)
        ctx.eng._sync_stream()
        ctx.eng.bind(self.data)
        self.saved = None

    def _vp(self, t):
        return self.ct.c_void_p(t.data_ptr())

    def bind(self):
        self.ctx.eng._sync_stream()
        self.ctx.eng.bind(self.data)

    def m_step(self):
        from mtm_b200 import _capi as C

        e = self.ctx.eng
        C.check(e.lib.mtm_mstep(e.h, self._vp(self.assign), self.n_c, 0.0, self._vp(self.params), None), e.h)

    def phase(self, which, info=None):
        """mtm_em_step: phase 1 = E-step kernel with the M-step statistics fused into its epilogue + the two
        small reduction kernels (+ the all-reduce when N > 1); phase 2 = parameter solve; 3 = both."""
        from mtm_b200 import _capi as C

        e = self.ctx.eng
        C.check(e.lib.mtm_em_step(e.h, self.n_c, 0.0, self._vp(self.assign), self._vp(self.params), which,
                                  self.ct.byref(info) if info is not None else None), e.h)

    def save(self):
        self.saved = (self.assign.clone(), self.params.clone())

    def restore(self):
        self.assign.copy_(self.saved[0])
        self.params.copy_(self.saved[1])

    def timed_steps(self, steps):
        """K stationary EM iterations; returns (ms_total, ms_e_mean, ms_m_mean), each the max over ranks."""
        ctx = self.ctx
        ev = [[ctx.event() for _ in range(3)] for _ in range(steps)]
        t_beg, t_end = ctx.event(), ctx.event()
        ctx.barrier()
        t_beg.record()
        for k in range(steps):
            self.restore()
            ev[k][0].record()
            self.phase(1)
            ev[k][1].record()
            self.phase(2)
            ev[k][2].record()
        t_end.record()
        ctx.barrier()
        ms_total = t_beg.elapsed_time(t_end)
        ms_e = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
        ms_m = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))
        return ctx.max_over_ranks([ms_total, ms_e, ms_m])

    def check(self, n_total):
        """What the timed iteration did, over all ranks: switches, cluster sizes, guard, status."""
        from mtm_b200 import _capi as C

        ctx, torch = self.ctx, self.ctx.torch
        info = C.EStepInfo()
        self.restore()
        self.phase(3, info)
        sizes = ctx.sum_over_ranks(torch.bincount(self.assign.long(), minlength=self.n_c).double())
        sw = ctx.sum_over_ranks((self.assign != self.saved[0]).sum().double().reshape(1))
        sizes = [int(v) for v in sizes.cpu()]
        st = ctx.sum_over_ranks(torch.tensor([float(info.status)], dtype=torch.float64, device=ctx.device))
        ok = min(sizes) > 3 and int(st.item()) == 0 and sum(sizes) == n_total
        return {"ok": bool(ok), "min_count": min(sizes), "min_count_library": int(info.min_count),
                "status": int(info.status), "cluster_sizes": sizes, "n_switches": int(sw.item()),
                "stats_mode": int(info.stats_mode)}


def em_roofline(ctx, cfg, ds, n_loc, ms_e, ms_m, pattern_order, kernel, wsc, delta=False):
    T, d, l, n_c = cfg["T"], cfg["d"], cfg["l"], cfg["n_c"]
    f_eval = flops_per_eval(T, d, l, ds["kz"], ds["kx"])
    f_survey = flops_per_eval_survey(T, d, l, ds["kz"] + ds["kx"])
    instr_eval = T * (162.0 * wsc["point"] + 264.0 * wsc["rank_one"] + 466.0 * wsc["generic"])
    evals = n_loc * n_c
    tf = f_eval * evals / (ms_e * 1e-3) * 1e-12
    bpt = bytes_per_traj(T, d, l)
    # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of THIS launch configuration
    # (profiles/r2i_fused_ncu_summary.md: 5.168793 GB read = one pass over the data + labels, 112.1 MB written = new
    # labels + 72 MB of block partials); any other size has no capture
    captured = (n_loc == 10_000_000 and (T, d, l, n_c) == (8, 4, 4, 8) and pattern_order and not delta)
    return {"bound": "fp64", "achieved": tf, "peak": ctx.fp64_peak, "unit": "TFLOP/s", "frac": tf / ctx.fp64_peak,
            "traffic": (5.168793e9 + 112.147968e6) if captured else None,
            "traffic_unit": "bytes per launch (ncu capture of this configuration)" if captured else None,
            "traffic_note": ("no ncu capture of the final delta twin (GPU budget of the round); its static-striding build "
                             "read 5.1626 GB and wrote 114.1 MB per launch at this size (profiles/r2_delta_ncu_summary.md), "
                             "of which 72 MB were block partials that the final one-wave grid cuts to 6 MB")
            if (delta and n_loc == 10_000_000 and (T, d, l, n_c) == (8, 4, 4, 8) and pattern_order) else None,
            "algorithmic_bytes": bpt * n_loc, "kernel": kernel, "kernel_ms": ms_e, "flops_per_eval": f_eval,
            "evals_per_launch": evals,
            # What the kernel EXECUTES: FP64 instructions per evaluation by code path (SASS counts of the time loop:
            # point-mass step 162, rank-one step 264, generic step 466) weighted with the measured path mix of the
            # resident order, and the FP64-pipe activity that implies (a warp instruction occupies an SM
            # sub-partition's FP64 pipe for 2 cycles; ncu measured 57.5 % for this kernel, profiles/)
            "fp64_instr_per_eval_model": instr_eval,
            "fp64_pipe_active_model": instr_eval * evals / 32.0 * 2.0 / (148 * 4 * 1.965e9 * ms_e * 1e-3),
            "warp_step_paths": wsc,
            "peak_source": "measured here: mtm_fp64_peak, the better of two DFMA micro-kernels, best of 10 each "
                           "(nominal 148 SM x 64 DFMA/clk x 2 x 1.965 GHz = 37.2; MEASURED_PEAKS.json has no fp64 entry)",
            "frac_of_nominal": tf / 37.2,
            "achieved_survey_formula": f_survey * evals / (ms_e * 1e-3) * 1e-12,
            "hbm": {"achieved": bpt * n_loc / (ms_e * 1e-3) * 1e-9, "peak": ctx.hbm_peak, "unit": "GB/s",
                    "frac": bpt * n_loc / (ms_e * 1e-3) * 1e-9 / ctx.hbm_peak, "bytes_per_trajectory": bpt,
                    "peak_source": ctx.hbm_src},
            "finalize_ms": ms_m}


# -------------------------------------------------------------------------------------------------
# the legs
# -------------------------------------------------------------------------------------------------
def leg_em(ctx, cfg, steps, warmup, tag, pattern_order=True):
    """EM iteration on a global synthetic set sharded over the ranks (C3 headline, C2)."""
    torch = ctx.torch
    n_total, T, d, l, n_c = cfg["n"], cfg["T"], cfg["d"], cfg["l"], cfg["n_c"]
    lo, hi = shard_bounds(n_total, ctx.world, ctx.rank)
    n_loc = hi - lo
    comps = mixture_params(n_c, d, l)
    z, x = sample_global(n_total, lo, hi, T, d, l, comps, ctx.device)
    a0 = labels_global(n_total, lo, hi, n_c, ctx.device)
    ds = data_stats(torch, z, x)
    run = EmRunner(ctx, z, x, a0, n_c, pattern_order=pattern_order)
    wsc = warp_step_classes(torch, run.data.z)
    run.m_step()  # parameters of the initial assignment (train()'s first M-step)
    for _ in range(warmup):
        run.phase(3)
    run.save()
    run.restore()
    run.phase(3)  # one untimed pass of the very iteration that is timed
    ctx.barrier()
    if ctx.fp64_peak is None:
        ctx.fp64_peak = ctx.eng.fp64_peak(10)  # TFLOP/s, own DFMA micro-kernel
        ctx.barrier()
    launches0 = ctx.eng.launch_count()
    with ClockSampler(ctx.local_rank) as clk:
        ms_total, ms_e, ms_m = run.timed_steps(steps)
    launches = ctx.eng.launch_count() - launches0
    ms_step = ms_total / steps
    chk = run.check(n_total)
    # how the M-step statistics of the timed iteration were formed (mtm_b200.h, "delta statistics": late iterations
    # move few trajectories and only those are visited), and the same iteration with the mode switched off
    k_full = max(1, min(steps, 5))
    assign_default, params_default = run.assign.clone(), run.params.clone()  # the timed iteration as shipped
    ctx.eng.set_delta_stats(False)
    run.restore()
    run.phase(3)
    full_total, full_e, _ = run.timed_steps(k_full)
    chk_full = run.check(n_total)
    params_full = run.params.clone()
    labels_same = bool(ctx.torch.equal(run.assign, assign_default))
    ctx.eng.set_delta_stats(True)  # (drops the cache)
    run.assign.copy_(assign_default)  # the runner goes on in the state the default iteration left
    run.params.copy_(params_default)
    statistics = {"mode_of_timed_iteration": "delta" if chk["stats_mode"] == 1 else "full",
                  "switch_fraction": chk["n_switches"] / n_total,
                  "ms_per_step_full_statistics": full_total / k_full, "kernel_ms_full_statistics": full_e,
                  "steps_full_statistics": k_full, "mode_when_switched_off": chk_full["stats_mode"],
                  "labels_equal_delta_vs_full": labels_same,
                  "params_max_rel_err_delta_vs_full": rel_err(params_default.cpu().numpy(), params_full.cpu().numpy()),
                  "note": "value / ms_per_step are the iteration as mtm_em_step and mtm_em_fit run it (delta statistics "
                          "when the previous iteration moved at most n/8 labels); the *_full_statistics fields time the "
                          "same iteration with every cluster re-summed from all members, as the reference's M_step does"}
    out = {"tag": tag, "n_total": n_total, "n_per_gpu": n_loc, "n_c": n_c, "steps": steps, "ms_per_step": ms_step,
           "value": n_total * n_c / (ms_step * 1e-3), "em_iters_per_s": 1e3 / ms_step, "clocks": clk.summary(),
           "gpu_launches": int(launches), "workload_check": chk, "statistics": statistics, "order_ms": run.order_ms,
           "present_scalars_per_trajectory": {"z": ds["kz"], "x": ds["kx"]}, "warp_step_paths": wsc,
           "roofline": em_roofline(ctx, cfg, ds, n_loc, ms_e, ms_m, pattern_order,
                                   ("estep_delta_kernel<4,4,CONSTP=1>: E-step + delta statistics of the movers"
                                    if chk["stats_mode"] == 1 else
                                    "estep_kernel<4,4,HAS_Z=1,CONSTP=1> with the M-step statistics fused in") +
                                   " (the timed span also holds the two label-hash passes, the two ~10 us reduction "
                                   "kernels and, for N > 1, the all-reduce)", wsc, delta=chk["stats_mode"] == 1)}
    # state after the timed iteration: labels in the caller's order, gathered; Q and LL all-reduced by the library
    labels = ctx.gather_labels(run.data.from_resident(run.assign), n_total)
    q, ll, st = ctx.eng.score(run.data, run.params, n_c)
    out["state"] = {"assign_sha256": hashlib.sha256(np.ascontiguousarray(labels, dtype=np.int32).tobytes()).hexdigest(),
                    "Q": q, "LL": ll, "score_status": st,
                    "params_abs_sum": float(run.params.abs().sum().item())}
    if ctx.rank == 0:
        out["parity"] = oracle_gate(ctx, z, x, run.params, n_c, T)
    out["_params_final"] = run.params.clone()
    return out, run, (z, x), labels


def replay_single_gpu(ctx, cfg, warmup, params_multi, labels_multi, pattern_order=True):
    """N > 1, rank 0, after the communicator is gone: the same W + 1 EM iterations on the WHOLE data set on one
    GPU.  The N-GPU run must have produced the same labels and (to 1e-9 relative) the same parameters."""
    n_total, T, d, l, n_c = cfg["n"], cfg["T"], cfg["d"], cfg["l"], cfg["n_c"]
    comps = mixture_params(n_c, d, l)
    z, x = sample_global(n_total, 0, n_total, T, d, l, comps, ctx.device)
    a0 = labels_global(n_total, 0, n_total, n_c, ctx.device)
    run = EmRunner(ctx, z, x, a0, n_c, pattern_order=pattern_order)
    del z, x
    run.m_step()
    for _ in range(warmup + 1):
        run.phase(3)
    ctx.torch.cuda.synchronize(ctx.device)
    lab1 = run.data.from_resident(run.assign).cpu().numpy()
    P1, PN = run.params.cpu().numpy(), params_multi.cpu().numpy()
    return {"against": "the same W + 1 iterations of the whole 10 M set on one GPU (rank 0, no communicator)",
            "assignments_equal": bool(np.array_equal(lab1, labels_multi)),
            "n_label_differences": int(np.sum(lab1 != labels_multi)),
            "params_max_rel_err": rel_err(PN, P1),
            "assign_sha256_single_gpu": hashlib.sha256(np.ascontiguousarray(lab1, dtype=np.int32).tobytes()).hexdigest()}


def leg_e2e(ctx, cfg, run, z, x, steps):
    """End to end through the host-buffer entry points: (a) per step — every step uploads the rank's whole shard
    from pinned host memory and downloads labels + parameters (mtm_em_step_host); (b) per fit — what train()
    does: upload once, K resident iterations, download; plus the plain concurrent H2D rate of the box."""
    torch = ctx.torch
    n_total, n_c = cfg["n"], cfg["n_c"]
    n_loc = x.shape[1]
    zh = torch.empty(z.shape, dtype=torch.float64, pin_memory=True)
    xh = torch.empty(x.shape, dtype=torch.float64, pin_memory=True)
    zh.copy_(z)
    xh.copy_(x)
    ah = torch.empty(n_loc, dtype=torch.int32, pin_memory=True)
    ph = torch.empty(run.params.shape, dtype=torch.float64, pin_memory=True)
    a_caller = run.data.from_resident(run.saved[0])
    params = run.saved[1].clone()
    assign_dev = torch.empty(n_loc, dtype=torch.int32, device=ctx.device)
    e2e_steps = max(1, min(steps, 3))

    def step():
        # public host-buffer entry point (mtm_em_step_host): H2D of this step's inputs from pinned memory in
        # slabs overlapped with the compute, D2H of the labels and parameters.  The few e2e steps continue the EM
        # from the saved state (restoring 40 MB of pinned labels per step would put host memcpy time into the
        # PCIe-bound number).
        ctx.eng.em_step_host(zh, xh, z, x, ah, assign_dev, params, ph, n_c, 0.0, n_slabs=16)

    ah.copy_(a_caller)
    step()
    ctx.barrier()
    t0, t1 = ctx.event(), ctx.event()
    t0.record()
    for _ in range(e2e_steps):
        step()
    t1.record()
    ctx.barrier()
    (ms,) = ctx.max_over_ranks([t0.elapsed_time(t1) / e2e_steps])
    out = {"value": n_total * n_c / (ms * 1e-3), "unit": UNIT,
           "h2d_bytes_per_step": int(zh.numel() * 8 + xh.numel() * 8 + ah.numel() * 4),
           "d2h_bytes_per_step": int(ah.numel() * 4 + ph.numel() * 8), "ms_per_step": ms, "steps": e2e_steps,
           "h2d_note": "z + x + old labels",
           "note": "per rank; every step = one mtm_em_step_host call: the rank's whole shard is uploaded from "
                   "pinned host memory in 16 slabs overlapped with the fused E-step kernel, labels and "
                   "parameters are downloaded"}
    # plain H2D rate, all ranks at once: what the box gives the per-step number above
    ctx.barrier()
    t0, t1 = ctx.event(), ctx.event()
    t0.record()
    z.copy_(zh, non_blocking=True)
    x.copy_(xh, non_blocking=True)
    t1.record()
    ctx.barrier()
    gbs = (zh.numel() + xh.numel()) * 8 / (t0.elapsed_time(t1) * 1e-3) * 1e-9
    lo_hi = ctx.max_over_ranks([gbs, -gbs])
    out["h2d_probe"] = {"gbs_per_rank_max": lo_hi[0], "gbs_per_rank_min": -lo_hi[1],
                        "note": "one cudaMemcpyAsync of the pinned shard per array, all ranks concurrently"}
    # per fit: upload once (pinned -> HBM), pattern order, k_fit resident iterations, labels back.  Run twice, report
    # the second (the first may pay the caching allocator's cudaMallocs of two more copies of the shard); both are
    # in the line
    k_fit = 5
    for attempt in range(2):
        ctx.barrier()
        t0, t1, t2 = ctx.event(), ctx.event(), ctx.event()
        t0.record()
        dd = ctx.mtm.DeviceData(zh, xh, ctx.device, pattern_order=True)
        aa = dd.to_resident(a_caller)
        t1.record()
        _, info = ctx.eng.em_fit(dd, aa, n_c, 0.0, k_fit)
        lab = dd.from_resident(aa).to("cpu", non_blocking=False)
        t2.record()
        ctx.barrier()
        ms_fit, ms_up = ctx.max_over_ranks([t0.elapsed_time(t2), t0.elapsed_time(t1)])
        if attempt == 0:
            first_ms = ms_fit
            ctx.eng.bind(run.data)
            del dd, aa
    out["per_fit"] = {"ms_total": ms_fit, "ms_upload_and_order": ms_up, "ms_total_first_call": first_ms,
                      "em_iterations": int(info["n_e_steps"]),
                      "exit": info["exit"], "value": n_total * n_c * max(1, int(info["n_e_steps"])) / (ms_fit * 1e-3),
                      "unit": UNIT, "labels_downloaded": int(lab.numel()),
                      "note": "DeviceData(pinned host z, x) + mtm_em_fit + labels back: the data cross PCIe once per fit, as in "
                              "train(); second of two identical calls"}
    ctx.eng.bind(run.data)
    del zh, xh, dd, aa
    return out


def leg_c4(ctx, args, params8, steps):
    """C4: measurement-only predict (hidden states marginalised, MMM:1125-1165 / XVAL:223-233) of held-out
    trajectories with the C3 model: no hidden states are generated into HBM or read, no collective."""
    torch = ctx.torch
    n_total, T, d, l, n_c = args.c4_trajectories, args.T, args.d, args.l, params8.shape[0]
    lo, hi = shard_bounds(n_total, ctx.world, ctx.rank)
    n_loc = hi - lo
    comps = mixture_params(n_c, d, l)
    _, x = sample_global(n_total, lo, hi, T, d, l, comps, ctx.device, seed=987_000, want_z=False)
    ds = data_stats(torch, None, x)
    data = ctx.mtm.DeviceData(None, x, ctx.device)
    assign = torch.empty(n_loc, dtype=torch.int32, device=ctx.device)
    call = lambda: ctx.eng.estep(data, params8, n_c, T, assign=assign, d=d, want_info=False)
    for _ in range(2):
        call()
    ev = [ctx.event() for _ in range(steps + 1)]
    with ClockSampler(ctx.local_rank) as clk:
        ctx.barrier()
        ev[0].record()
        for k in range(steps):
            call()
            ev[k + 1].record()
        ctx.barrier()
    (ms,) = ctx.max_over_ranks([ev[0].elapsed_time(ev[steps]) / steps])
    f_eval = flops_per_eval(T, d, l, 0.0, ds["kx"])
    tf = f_eval * n_loc * n_c / (ms * 1e-3) * 1e-12
    bpt = 8 * T * l + 4
    out = {"workload": f"C4: observation-only predict of {n_total / 1e6:g}M held-out trajectories, n_c={n_c}, hidden "
                       f"states marginalised, sharded x{ctx.world}", "n_total": n_total, "n_per_gpu": n_loc,
           "steps": steps, "ms": ms, "evals_per_s": n_total * n_c / (ms * 1e-3), "clocks": clk.summary(),
           "roofline": {"bound": "fp64", "achieved": tf, "peak": ctx.fp64_peak, "unit": "TFLOP/s",
                        "frac": tf / ctx.fp64_peak, "flops_per_eval": f_eval, "kernel": "estep_kernel<4,4,HAS_Z=0,CONSTP=1>",
                        "hbm": {"achieved": bpt * n_loc / (ms * 1e-3) * 1e-9, "peak": ctx.hbm_peak, "unit": "GB/s",
                                "bytes_per_trajectory": bpt}},
           "cluster_sizes": [int(v) for v in ctx.sum_over_ranks(torch.bincount(assign.long(), minlength=n_c).double()).cpu()]}
    if ctx.rank == 0:
        out["parity"] = oracle_gate(ctx, None, x, params8, n_c, T, d=d)
    del x, data
    return out


def leg_c5(ctx, cfg, run, z, x, n_restarts, ncs, n_steps=2):
    """C5: model-selection sweep (MSEL:45-54): for every n_c, `n_restarts` random restarts advance together through
    mtm_sweep_fit (batched restarts, kernel 4) on the resident C3 set."""
    torch = ctx.torch
    n_total, T, d, l = cfg["n"], cfg["T"], cfg["d"], cfg["l"]
    lo, hi = shard_bounds(n_total, ctx.world, ctx.rank)
    n_loc = hi - lo
    ds = data_stats(torch, z, x)
    sets, ms_sum, per_nc, exits = 0, 0.0, {}, {}
    # untimed: one iteration at the largest n_c sizes the library's scratch and the caching allocator's blocks once
    # (a cudaFree + cudaMalloc of a grown buffer inside a timed call costs 50-250 ms with peers mapped)
    a = run.data.to_resident(labels_global(n_total, lo, hi, max(ncs), ctx.device, seed=7, rows=n_restarts))
    ctx.eng.sweep_fit(run.data, a, max(ncs), 0.0, 1, want_scores=False)
    del a
    with ClockSampler(ctx.local_rank) as clk:
        for n_c in ncs:
            a = labels_global(n_total, lo, hi, n_c, ctx.device, seed=100_000 * n_c, rows=n_restarts)
            a = run.data.to_resident(a)
            ctx.barrier()
            t0, t1 = ctx.event(), ctx.event()
            t0.record()
            _, infos, _, _ = ctx.eng.sweep_fit(run.data, a, n_c, 0.0, n_steps, want_scores=False)
            t1.record()
            ctx.barrier()
            (ms,) = ctx.max_over_ranks([t0.elapsed_time(t1)])
            per_nc[str(n_c)] = ms / n_steps
            ms_sum += ms
            sets += n_restarts * n_c
            for i in infos:
                exits[i["exit"]] = exits.get(i["exit"], 0) + 1
            del a
    sec_iter = ms_sum * 1e-3 / n_steps
    f_eval = flops_per_eval(T, d, l, ds["kz"], ds["kx"])
    tf = f_eval * sets * n_loc / sec_iter * 1e-12
    out = {"workload": f"C5: restart sweep n_c={ncs[0]}..{ncs[-1]} x {n_restarts} restarts on {n_total / 1e6:g}M "
                       f"trajectories, sharded x{ctx.world}", "parameter_sets": sets, "em_iterations": n_steps,
           "seconds_per_sweep_iteration": sec_iter, "evals_per_s": sets * n_total / sec_iter,
           "ms_per_iteration_by_n_c": per_nc, "exits": exits, "clocks": clk.summary(),
           "roofline": {"bound": "fp64", "achieved": tf, "peak": ctx.fp64_peak, "unit": "TFLOP/s",
                        "frac": tf / ctx.fp64_peak, "flops_per_eval": f_eval,
                        "note": "whole-call time (each call also runs the first M-step of its restarts and the "
                                "reductions / all-reduce / solves of every iteration), not the kernel alone"},
           "note": "each sweep_fit call = guard + first M-step + n_steps x (E + M) for all restarts of one n_c"}
    if ctx.rank == 0:
        out["parity"] = sweep_gate(ctx, z, x)
    ctx.eng.bind(run.data)
    return out


def leg_c1(ctx, n_starts=1000):
    """C1: the reference's own CPU-runnable case — fold 0 of inference-adni-xval.py (ADNI: n = 513, T = 4, d = 2, l = 4,
    n_c = 4, alpha = 1; XVAL:24-25, 127-133): train_with_multiple_random_starts(n_starts=1000) through the drop-in class
    (k-means start + 1000 batched random restarts), rank 0 only.  Gate: with the golden file's number of starts the
    objectives, the winner and its labels are the reference's."""
    from mtm_b200.framework import marginalizable_mixture_model as mixmodel

    g = np.load(os.path.join(ROOT, "tests", "golden", "adni_fold0.npz"))
    ztr, xtr, alpha = g["ztr"], g["xtr"], float(g["alpha"])
    make = lambda: mixmodel.MMLinGaussSS_marginalizable(n_clusters=4, states=ztr, observations=xtr, init="k-means",
                                                        alpha=alpha, device=ctx.device, group=False)
    best, obj = make().train_with_multiple_random_starts(n_starts=int(g["n_starts"]), use_cache=False, return_objectives=True)
    gate = {"objectives_max_rel_err": rel_err(obj, g["objectives"]), "winner_seed_equal": best.random_seed == int(g["best_seed"]),
            "winner_labels_equal": bool(np.array_equal(best.cluster_assignment, g["best_assign"]))}
    ctx.torch.cuda.synchronize(ctx.device)
    t0 = time.perf_counter()
    best = make().train_with_multiple_random_starts(n_starts=n_starts, use_cache=False)
    ctx.torch.cuda.synchronize(ctx.device)
    dt = time.perf_counter() - t0
    return {"workload": f"C1: ADNI fold 0 (n={ztr.shape[1]}, T=4, d=2, l=4, n_c=4, alpha=1), k-means start + {n_starts} random "
                        "restarts, one GPU", "seconds": dt, "fits_per_s": (n_starts + 1) / dt, "best_seed": int(best.random_seed),
            "best_Q": float(best.e_complete_data_log_lik()),
            "note": "host wall clock around the class call (latency-bound: 513 trajectories); the reference needs ~0.68 s per "
                    "fit on one core (SURVEY 6)", "parity": gate}


def sweep_gate(ctx, z, x, npre=4096, n_c=3, R=4, n_steps=3):
    """Batched restarts on a caller-order prefix against sequential fits of the CPU oracle: switch histories,
    final labels and Q of every restart."""
    from oracle import cport as oc
    from oracle import dense as od

    torch = ctx.torch
    zs, xs = z[:, :npre].contiguous(), x[:, :npre].contiguous()
    data = ctx.mtm.DeviceData(zs, xs, ctx.device)
    a0 = np.stack([np.random.default_rng(100 + i).integers(0, n_c, npre) for i in range(R)]).astype(np.int32)
    a = torch.from_numpy(a0).to(ctx.device)
    # a communicator may be attached: the prefix lives on rank 0 only, so fit it with a private handle
    eng = ctx.mtm.engine.Engine(ctx.device) if ctx.world > 1 else ctx.eng
    P, infos, scores, _ = eng.sweep_fit(data, a, n_c, 0.0, n_steps)
    zh, xh = zs.cpu().numpy(), xs.cpu().numpy()
    ok, qerr = True, 0.0
    for i in range(R):
        ref = od.train(zh, xh, a0[i].astype(np.int64), n_c, alpha=0.0, n_steps=n_steps, e_step_fn=oc.e_step_seq)
        ok = ok and infos[i]["switches"] == ref.switches and np.array_equal(a[i].cpu().numpy(), ref.assign)
        ll = oc.seq_loglik_matrix(ref.params.packed(), zh, xh)
        new = oc.assign(ll, ref.params.pi)
        q_ref = float(np.sum(np.log(ref.params.pi)[new] + ll[new, np.arange(npre)]))
        qerr = max(qerr, abs(scores[i, 0] - q_ref) / abs(q_ref))
    return {"trajectories": npre, "restarts": R, "n_c": n_c, "em_steps": n_steps,
            "switch_histories_and_labels_equal": bool(ok), "max_rel_Q_err": qerr}


# -------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--trajectories", "--n", dest="n", type=int, default=10_000_000,
                    help="total trajectories (config C3: 10 M); use --trajectories under torchrun (its parser grabs --n)")
    ap.add_argument("--n-clusters", type=int, default=8)
    ap.add_argument("--T", type=int, default=8)
    ap.add_argument("--d", type=int, default=4)
    ap.add_argument("--l", type=int, default=4)
    ap.add_argument("--cpu-sample", type=int, default=20000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--legs", default="c1,c2,c4,c5", help="other BASELINE configurations to measure (comma list, '' = none)")
    ap.add_argument("--c4-trajectories", type=int, default=100_000_000)
    ap.add_argument("--c5-restarts", type=int, default=32)
    ap.add_argument("--no-replay", action="store_true", help="N > 1: skip the single-GPU replay of the timed fit")
    ap.add_argument("--no-pattern-order", action="store_true",
                    help="keep the resident set in the caller's order (no point-mass shortcut for whole warps)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3  # timing rule: at least 3 warm-up steps

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = dict(n=args.n, T=args.T, d=args.d, l=args.l, n_c=args.n_clusters)
    workload = (f"C3: synthetic {args.n / 1e6:g}M trajectories, T={args.T}, d={args.d}, l={args.l}, "
                f"n_c={args.n_clusters}, 20% element-wise missing, one EM iteration per step")

    # ---------------- reference arm: CPU, rank 0 only ----------------
    if args.impl == "reference":
        if rank != 0:
            return
        base = run_cpu_arm(args, cfg, args.steps, args.warmup)
        line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["ms_per_step"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": workload, "sample": base["sample"]},
                "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "em_iters_per_s_on_sample": 1e3 / base["ms_per_step"]}
        print(json.dumps(line))
        return

    # ---------------- B200 arm ----------------
    # stdout carries exactly one JSON line: anything a library prints on fd 1 meanwhile (NCCL's version banner at
    # communicator creation) goes to stderr; fd 1 is restored right before the line is printed
    sys.stdout.flush()
    stdout_fd = os.dup(1)
    os.dup2(2, 1)
    ctx = Ctx(args)
    torch = ctx.torch
    legs = [s for s in args.legs.lower().split(",") if s]
    po = not args.no_pattern_order

    main_leg, run, (z, x), labels = leg_em(ctx, cfg, args.steps, args.warmup, "C3", pattern_order=po)
    params_final = main_leg.pop("_params_final")
    e2e = None if args.no_e2e else leg_e2e(ctx, cfg, run, z, x, args.steps)
    params_model = run.saved[1].clone()  # the model after the warm-up iterations: what C4 predicts with

    configs = {}
    if "c5" in legs and cfg["n_c"] >= 1:
        configs["C5"] = leg_c5(ctx, cfg, run, z, x, args.c5_restarts, list(range(2, 13)))
    del run, z, x
    torch.cuda.empty_cache()
    if "c2" in legs:
        c2cfg = dict(n=1_000_000, T=8, d=4, l=4, n_c=4)
        # always 3 warm-up iterations: on this small 4-cluster set a cluster keeps shrinking, the timed state
        # should not depend on the driver's --warmup
        c2, run2, zx2, _ = leg_em(ctx, c2cfg, max(args.steps, 20), 3, "C2", pattern_order=po)
        c2["workload"] = "C2: synthetic 1M trajectories, T=8, d=l=4, n_c=4, 20% missing, one EM iteration per step"
        c2["evals_per_s"] = c2.pop("value")
        c2.pop("_params_final")
        configs["C2"] = c2
        del run2, zx2
        torch.cuda.empty_cache()
    if "c4" in legs:
        configs["C4"] = leg_c4(ctx, args, params_model, max(3, min(args.steps, 5)))
        torch.cuda.empty_cache()
    if "c1" in legs and rank == 0:
        configs["C1"] = leg_c1(ctx)

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:  # reported at N = 1 only
        b = run_cpu_arm(args, cfg, steps=1, warmup=0)
        cpu_base = {k: b[k] for k in ("value", "unit", "cores", "kind", "sample")}

    class_gate = None
    if world > 1:
        # the drop-in class sharding itself over the same process group (its own handle and communicator)
        from mtm_b200.framework import marginalizable_mixture_model as mixmodel
        from tools.class_dist_checks import class_checks

        failures = class_checks(mixmodel, world, rank)
        class_gate = {"ok": not failures, "failures": failures[:5],
                      "what": "MMLinGaussSS_marginalizable built from the full arrays on every rank: golden traces c2s / "
                              "adnish / guardexit, held-out and observation-only predictions, restart selection, "
                              "communicator re-attach"}
        ctx.eng.comm_destroy()
    multi = None
    if world > 1 and rank == 0 and not args.no_replay:
        multi = replay_single_gpu(ctx, cfg, args.warmup, params_final, labels, pattern_order=po)
    if world > 1:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()

    rc = 0
    if rank == 0:
        parity = main_leg.pop("parity")
        parity["multi_gpu"] = multi
        parity["class_sharded"] = class_gate
        chk = main_leg["workload_check"]
        line = {"metric": METRIC, "value": main_leg["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": main_leg["ms_per_step"], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "n_total": cfg["n"], "n_per_gpu": main_leg["n_per_gpu"],
                           "data_generation": f"{G_CHUNKS} seeded chunks of the global trajectory axis: identical data "
                                              "and initial labels at every GPU count",
                           "stationarity": f"state after {args.warmup} warm-up EM iterations saved and restored (2 "
                                           "device-to-device copies, inside the timed region) before every timed step",
                           "residency": ("caller's order" if args.no_pattern_order else
                                         f"pattern-ordered (mtm_pattern_order + row permutation, one-off "
                                         f"{main_leg['order_ms']:.1f} ms when the set is made resident; outside the timed steps)"),
                           "l2_policy": "inputs (512 B x n per GPU) far larger than the 126 MB L2; no flush needed",
                           "parallelism": f"trajectory shards x{world}, one all-reduce of "
                                          f"{cfg['n_c'] * ctx.eng.stats_stride(args.d, args.l) + 4} doubles per iteration "
                                          f"({os.environ.get('MTM_B200_ALLREDUCE', 'p2p') if world > 1 else 'none'})"},
                "em_iters_per_s": main_leg["em_iters_per_s"],
                "clocks": main_leg["clocks"], "e2e": e2e, "gpu_launches": main_leg["gpu_launches"],
                "roofline": main_leg["roofline"], "cpu_baseline": cpu_base, "parity": parity,
                "workload_check": chk, "statistics": main_leg["statistics"], "state": main_leg["state"],
                "present_scalars_per_trajectory": main_leg["present_scalars_per_trajectory"],
                "warp_step_paths": main_leg["warp_step_paths"], "configs": configs}
        bad = []
        if not chk["ok"]:
            bad.append(f"timed EM iteration is degenerate: {chk}")
        if not (parity["assignments_equal"] and parity["max_rel_ll_err"] < 1e-9):
            bad.append("oracle parity gate failed")
        if multi is not None and not (multi["assignments_equal"] and multi["params_max_rel_err"] < 1e-9):
            bad.append("N-GPU fit differs from the single-GPU fit")
        if class_gate is not None and not class_gate["ok"]:
            bad.append("sharded drop-in class differs from the reference's golden traces")
        if bad:
            line["invalid"] = bad
            rc = 1
    sys.stdout.flush()
    os.dup2(stdout_fd, 1)
    if rank == 0:
        print(json.dumps(line), flush=True)
    sys.exit(rc)


if __name__ == "__main__":
    main()
