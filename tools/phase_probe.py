#!/usr/bin/env python3
"""Probe: where the EM iteration's time goes on the pattern-ordered C3 set: bare E-step (mtm_estep, no
statistics), fused E-step + statistics (mtm_em_step phase 1), stand-alone statistics (mtm_mstats), and the
observation-only kernel on the same x."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import mtm_b200  # noqa: E402
from mtm_b200 import _capi as C  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
T, d, l, n_c = 8, 4, 4, 8
dev = torch.device("cuda", 0)
eng = mtm_b200.Engine.get(dev)
comps = bench.mixture_params(n_c, d, l)
z, x = bench.sample_device(n, T, d, l, comps, seed=4321, device=dev)
data = mtm_b200.DeviceData(z, x, dev, pattern_order=True)
del z, x
assign = torch.randint(0, n_c, (n,), device=dev, dtype=torch.int32)
vp = lambda t: ctypes.c_void_p(t.data_ptr())
eng._sync_stream()
eng.bind(data)
params = torch.empty((n_c, eng.param_stride(d, l)), dtype=torch.float64, device=dev)
C.check(eng.lib.mtm_mstep(eng.h, vp(assign), n_c, 0.0, vp(params), None), eng.h)
for _ in range(2):
    C.check(eng.lib.mtm_em_step(eng.h, n_c, 0.0, vp(assign), vp(params), 3, None), eng.h)
obs = mtm_b200.DeviceData.__new__(mtm_b200.DeviceData)
obs.z, obs.x, obs.T, obs.n, obs.l, obs.d, obs.perm = None, data.x, T, n, l, None, data.perm
stats = torch.empty((n_c, eng.stats_stride(d, l)), dtype=torch.float64, device=dev)


def timed(tag, fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        fn()
    t1.record()
    torch.cuda.synchronize()
    print(f"{tag:46s} {t0.elapsed_time(t1) / reps:8.3f} ms", flush=True)


a2 = assign.clone()
timed("bare E-step (labels + switch count)", lambda: eng.estep(data, params, n_c, T, assign=a2, count_switches=True, want_info=False))
timed("fused E-step + statistics (em_step phase 1)", lambda: C.check(eng.lib.mtm_em_step(eng.h, n_c, 0.0, vp(assign), vp(params), 1, None), eng.h))
eng.bind(data)
timed("stand-alone statistics kernel (mtm_mstats)", lambda: C.check(eng.lib.mtm_mstats(eng.h, vp(assign), n_c, vp(stats)), eng.h))
timed("observation-only E-step on the same x", lambda: eng.estep(obs, params, n_c, T, assign=a2, d=d, want_info=False))
