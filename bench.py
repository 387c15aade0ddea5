#!/usr/bin/env python3
"""Benchmark of the hard-assignment EM hot path (BASELINE.json: trajectory x cluster log-lik
evaluations / s and EM iterations / s at 10 M trajectories).

    python bench.py --gpus N --steps K --warmup W            # this implementation (B200)
    python bench.py --impl reference --steps K --warmup W     # CPU port of the reference algorithm

A *step* is one EM iteration over the whole data set: E-step (n x n_c log-likelihood
evaluations + argmax + switch count) followed by the M-step (sufficient statistics, the
all-reduce when N > 1, parameter solve).  Exactly K steps are timed, with no convergence exit.
Workload at N = 1: config C3 of BASELINE.json (10 M trajectories, T = 8, d = l = 4, n_c = 8,
20 % element-wise missing); N > 1 shards those same 10 M trajectories (strong scaling).

One JSON line on stdout (rank 0).  See the module-level comments at each block for how every
number is obtained.
"""

from __future__ import annotations

import argparse
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


def flops_skipped_by_point_steps(T, d, l, n_point, n_point_not_last, kx_point):
    """Part of flops_per_eval the kernel does NOT execute when a warp takes the point-mass shortcut
    (all of z_t present for its 32 trajectories): the x-predictive covariance products and the (mu, P, C)
    updates of that step work on an exactly zero covariance, and so does A'PA of the following propagation.
    n_point = such steps per trajectory, kx_point = present x scalars in them."""
    pred_cov = 2 * d * d * l + d * l * (l + 1)                    # C = P H and H'C
    per_x = d + 2 * d + d * (d + 1) + 0.5 * (l - 1) * 2 * d       # gains, mu, P and C updates per present x scalar
    prop_cov = 2 * d ** 3 + d * d * (d + 1)                       # P A and A'(PA)
    return n_point * pred_cov + kx_point * per_x + n_point_not_last * prop_cov


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
def cpu_em_iteration(z, x, assign, n_c, par, nthreads):
    """One EM iteration the way the reference does it: dense joint covariance per cluster,
    per-row det + solve by LU (C port, OpenMP over rows), numpy M-step (oracle/dense.py)."""
    from oracle import cport as oc
    from oracle import dense as od

    ll = oc.dense_loglik_matrix(par, z, x, nthreads=nthreads)
    new = od.assign_from_loglik(par.pi, ll)
    return new, od.m_step(z, x, new, n_c, 0.0)


def run_cpu_arm(args, cfg, steps, warmup):
    from oracle import cport as oc
    from oracle import dense as od

    n_s = args.cpu_sample
    comps = mixture_params(cfg["n_c"], cfg["d"], cfg["l"])
    z, x = sample_host(n_s, cfg["T"], cfg["d"], cfg["l"], comps, seed=4321)
    assign = od.random_init(n_s, cfg["n_c"], 42)
    par = od.m_step(z, x, assign, cfg["n_c"], 0.0)
    threads = oc.num_threads()
    for _ in range(warmup):
        assign, par = cpu_em_iteration(z, x, assign, cfg["n_c"], par, threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        assign, par = cpu_em_iteration(z, x, assign, cfg["n_c"], par, threads)
    dt = time.perf_counter() - t0
    value = n_s * cfg["n_c"] * steps / dt
    return dict(value=value, unit=UNIT, cores=threads, kind="port",
                sample=f"{n_s} trajectories of the same synthetic workload x {steps} EM iterations "
                       f"(dense-covariance LU port of SSM:728-773 in C with OpenMP + numpy M-step)",
                ms_per_step=1e3 * dt / steps, host_cpus=os.cpu_count())


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
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--cpu-sample", type=int, default=20000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-pattern-order", action="store_true",
                    help="keep the resident set in the caller's order (no point-mass shortcut for whole warps)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3  # timing rule: at least 3 warm-up steps

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
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
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": workload, "sample": base["sample"]},
                "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "em_iters_per_s_on_sample": 1e3 / base["ms_per_step"]}
        print(json.dumps(line))
        return

    # ---------------- B200 arm ----------------
    import torch
    import torch.distributed as dist

    import mtm_b200
    from mtm_b200 import _capi as C

    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    # stdout carries exactly one JSON line: anything a library prints on fd 1 meanwhile (NCCL's version banner at
    # communicator creation) goes to stderr; fd 1 is restored right before the line is printed
    sys.stdout.flush()
    stdout_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    eng = mtm_b200.Engine.get(device)
    if world > 1:
        from mtm_b200 import dist as mdist

        mdist.attach(eng)  # library-owned NCCL communicator + peer-memory mailboxes (MTM_B200_ALLREDUCE=nccl|p2p)

    n_total = args.n if args.scaling == "strong" else args.n * world
    lo = rank * n_total // world
    hi = (rank + 1) * n_total // world
    n_loc = hi - lo
    comps = mixture_params(cfg["n_c"], cfg["d"], cfg["l"])
    z, x = sample_device(n_loc, args.T, args.d, args.l, comps, seed=4321 + rank, device=device)
    # resident order: trajectories grouped by their per-step "z complete" pattern (mtm_pattern_order, a one-off
    # pass when the training set is made resident, like the upload; timed separately, outside the steps)
    t_o0, t_o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_o0.record()
    data = mtm_b200.DeviceData(z, x, device, pattern_order=not args.no_pattern_order)
    t_o1.record()
    torch.cuda.synchronize(device)
    order_ms = t_o0.elapsed_time(t_o1)
    gen = torch.Generator(device=device)
    gen.manual_seed(42 + rank)
    assign = torch.randint(0, cfg["n_c"], (n_loc,), generator=gen, device=device, dtype=torch.int32)
    kz = float(torch.isfinite(z).sum().item()) / n_loc
    kx = float(torch.isfinite(x).sum().item()) / n_loc
    zc = torch.isfinite(z).all(dim=2)                                        # (T, n): z_t complete
    n_point = float(zc.sum().item()) / n_loc
    n_point_nl = float(zc[:-1].sum().item()) / n_loc
    kx_point = float((torch.isfinite(x).sum(dim=2) * zc).sum().item()) / n_loc
    del zc
    lib, h = eng.lib, eng.h
    import ctypes

    eng._sync_stream()
    eng.bind(data)
    st = data.struct()
    pstride = eng.param_stride(args.d, args.l)
    params = torch.empty((cfg["n_c"], pstride), dtype=torch.float64, device=device)
    vp = lambda t: ctypes.c_void_p(t.data_ptr())

    def m_step():
        C.check(lib.mtm_mstep(h, vp(assign), cfg["n_c"], 0.0, vp(params), None), h)

    # one EM iteration = mtm_em_step: phase 1 = E-step kernel with the M-step statistics fused into its
    # epilogue + the two small reduction kernels (+ the all-reduce when N > 1); phase 2 = parameter solve
    def e_step():
        C.check(lib.mtm_em_step(h, cfg["n_c"], 0.0, vp(assign), vp(params), 1, None), h)

    def m_finish():
        C.check(lib.mtm_em_step(h, cfg["n_c"], 0.0, vp(assign), vp(params), 2, None), h)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    m_step()  # parameters of the initial assignment (train()'s first M-step)
    for _ in range(args.warmup):
        e_step()
        m_finish()
    barrier()
    fp64_peak = eng.fp64_peak(10)  # TFLOP/s, own DFMA micro-kernel (MEASURED_PEAKS.json has no fp64 entry)
    barrier()

    # ---- timed region: exactly K EM iterations, inputs (n_loc * 512 B >> 126 MB L2) resident in HBM ----
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    t_beg, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = eng.launch_count()
    with ClockSampler(local_rank) as clk:
        barrier()
        t_beg.record()
        for k in range(args.steps):
            ev[k][0].record()
            e_step()
            ev[k][1].record()
            m_finish()
            ev[k][2].record()
        t_end.record()
        barrier()
    launches = eng.launch_count() - launches0
    ms_total = t_beg.elapsed_time(t_end)
    ms_e = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    ms_m = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))
    if world > 1:
        t = torch.tensor([ms_total, ms_e, ms_m], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, ms_e, ms_m = (float(v) for v in t.cpu())
    ms_step = ms_total / args.steps
    value = n_total * cfg["n_c"] / (ms_step * 1e-3)

    # ---- roofline of the dominant kernel (E-step), per launch, this rank's shard ----
    f_eval = flops_per_eval(args.T, args.d, args.l, kz, kx)
    f_survey = flops_per_eval_survey(args.T, args.d, args.l, kz + kx)
    f_skip = 0.0 if args.no_pattern_order else flops_skipped_by_point_steps(args.T, args.d, args.l, n_point,
                                                                          n_point_nl, kx_point)
    evals_launch = n_loc * cfg["n_c"]
    achieved_tf = f_eval * evals_launch / (ms_e * 1e-3) * 1e-12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    bpt = bytes_per_traj(args.T, args.d, args.l)
    roofline = {"bound": "fp64", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": achieved_tf / fp64_peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel
                # (profiles/r1s_fused_ncu_summary.md: 1.0349 GB + 76.2 MB for 2 M trajectories = 555.5 B per
                # trajectory: 517 B read = one pass over the data + labels, 38 B written = labels + block partials),
                # scaled to this launch's trajectory count
                "traffic": 555.5 * n_loc, "traffic_unit": "bytes per launch (ncu-measured 555.5 B/trajectory x n)",
                "kernel": "estep_kernel<4,4,HAS_Z,CONSTP,FUSE> (E-step + fused M-step statistics; the timed span also "
                          "holds the two ~10 us reduction kernels and, for N > 1, the all-reduce)",
                "kernel_ms": ms_e,
                "flops_per_eval": f_eval, "evals_per_launch": evals_launch,
                # the algorithmic count above is layout-independent; on the pattern-ordered resident set whole warps
                # skip the zero-covariance arithmetic of the steps with a complete z_t, so the kernel EXECUTES less:
                "flops_per_eval_executed": f_eval - f_skip,
                "frac_executed": (f_eval - f_skip) * evals_launch / (ms_e * 1e-3) * 1e-12 / fp64_peak,
                "point_steps_per_trajectory": n_point,
                "peak_source": "measured here: mtm_fp64_peak, the better of two DFMA micro-kernels, best of 10 each "
                               "(nominal 37.2; the same instruction stream as a stand-alone probe on an idle device: "
                               "36.0-36.9, profiles/r1p_mix_probe.txt)",
                "frac_of_nominal": achieved_tf / 37.2,
                "achieved_survey_formula": f_survey * evals_launch / (ms_e * 1e-3) * 1e-12,
                "hbm": {"achieved": bpt * n_loc / (ms_e * 1e-3) * 1e-9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": bpt * n_loc / (ms_e * 1e-3) * 1e-9 / hbm_peak, "bytes_per_trajectory": bpt,
                        "peak_source": hbm_src},
                "finalize_ms": ms_m}

    # ---- parity gate on a prefix, against the CPU oracle ----
    parity = None
    if rank == 0:
        from oracle import cport as oc

        npre = 4096
        sub = mtm_b200.DeviceData(z[:, :npre].contiguous(), x[:, :npre].contiguous(), device)
        r = eng.estep(sub, params, cfg["n_c"], args.T, want_ll=True)
        P = params.cpu().numpy()
        llo = oc.seq_loglik_matrix(P, sub.z.cpu().numpy(), sub.x.cpu().numpy())
        ll = r["ll"].cpu().numpy()
        with np.errstate(all="ignore"):
            err = float(np.nanmax(np.abs(ll - llo) / np.maximum(1.0, np.abs(llo))))
        with np.errstate(all="ignore"):
            pren = np.sort(np.log(P[:, :1]) + llo, axis=0)  # log pi_c + ll_c, ascending over clusters
            margin = float(np.nanmin(pren[-1] - pren[-2])) if cfg["n_c"] > 1 else None
        parity = {"trajectories": npre, "max_rel_ll_err": err,
                  "assignments_equal": bool(np.array_equal(r["assign"].cpu().numpy(), oc.assign(llo, P[:, 0]))),
                  # smallest gap between the best and the second-best cluster of a trajectory (log domain): how
                  # far the closest call in the sample is from a tie that rounding could flip
                  "min_top2_margin": margin}

    # ---- end to end: host buffers in, labels + parameters out, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        zh = torch.empty(z.shape, dtype=torch.float64, pin_memory=True)
        xh = torch.empty(x.shape, dtype=torch.float64, pin_memory=True)
        zh.copy_(z)
        xh.copy_(x)
        ah = torch.empty(n_loc, dtype=torch.int32, pin_memory=True)
        ph = torch.empty(params.shape, dtype=torch.float64, pin_memory=True)
        e2e_steps = max(1, min(args.steps, 3))

        def e2e_step():
            # public host-buffer entry point (mtm_em_step_host): H2D of this step's inputs from pinned
            # memory in slabs overlapped with the compute, D2H of the labels and parameters
            eng.em_step_host(zh, xh, z, x, ah, assign, params, ph, cfg["n_c"], 0.0, n_slabs=16)

        ah.copy_(assign)
        e2e_step()
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(e2e_steps):
            e2e_step()
        t1.record()
        barrier()
        ms = t0.elapsed_time(t1) / e2e_steps
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        e2e = {"value": n_total * cfg["n_c"] / (ms * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(zh.numel() * 8 + xh.numel() * 8 + ah.numel() * 4),
               "d2h_bytes_per_step": int(ah.numel() * 4 + ph.numel() * 8), "ms_per_step": ms, "steps": e2e_steps,
               "h2d_note": "z + x + old labels",
               "note": "per rank; every step = one mtm_em_step_host call: the rank's whole shard is uploaded from "
                       "pinned host memory in 16 slabs overlapped with the fused E-step kernel, labels and "
                       "parameters are downloaded"}

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:  # reported at N = 1 only
        b = run_cpu_arm(args, cfg, steps=1, warmup=0)
        cpu_base = {k: b[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling,
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "n_total": n_total, "n_per_gpu": n_loc,
                           "residency": ("caller's order" if args.no_pattern_order else
                                         f"pattern-ordered (mtm_pattern_order + row permutation, one-off {order_ms:.1f} ms "
                                         "when the set is made resident; outside the timed steps)"),
                           "l2_policy": "inputs (512 B x n per GPU) far larger than the 126 MB L2; no flush needed",
                           "parallelism": f"trajectory shards x{world}, one all-reduce of "
                                          f"{cfg['n_c'] * eng.stats_stride(args.d, args.l) + 2} doubles per iteration "
                                          f"({os.environ.get('MTM_B200_ALLREDUCE', 'p2p') if world > 1 else 'none'})"},
                "em_iters_per_s": 1e3 / ms_step,
                "clocks": clk.summary(), "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
                "cpu_baseline": cpu_base, "parity": parity,
                "present_scalars_per_trajectory": {"z": kz, "x": kx}}
    if world > 1:
        eng.comm_destroy()
        dist.destroy_process_group()
    sys.stdout.flush()
    os.dup2(stdout_fd, 1)
    if rank == 0:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
