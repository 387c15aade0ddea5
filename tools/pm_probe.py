#!/usr/bin/env python3
"""Probe: EM-iteration time of the library named by MTM_B200_LIB on the C3 workload, with the trajectories in
their generated order and ordered by the per-step "all hidden-state coordinates present" pattern.
usage: MTM_B200_LIB=... python tools/pm_probe.py [n]"""
import ctypes
import os
import sys

import numpy as np
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
gen = torch.Generator(device=dev)
gen.manual_seed(42)
assign0 = torch.randint(0, n_c, (n,), generator=gen, device=dev, dtype=torch.int32)
vp = lambda t: ctypes.c_void_p(t.data_ptr())


def run(tag, z, x, assign):
    data = mtm_b200.DeviceData(z, x, dev)
    eng._sync_stream()
    eng.bind(data)
    params = torch.empty((n_c, eng.param_stride(d, l)), dtype=torch.float64, device=dev)
    C.check(eng.lib.mtm_mstep(eng.h, vp(assign), n_c, 0.0, vp(params), None), eng.h)
    for _ in range(3):
        C.check(eng.lib.mtm_em_step(eng.h, n_c, 0.0, vp(assign), vp(params), 3, None), eng.h)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    K = 5
    for _ in range(K):
        C.check(eng.lib.mtm_em_step(eng.h, n_c, 0.0, vp(assign), vp(params), 3, None), eng.h)
    t1.record()
    torch.cuda.synchronize()
    print(f"{os.environ.get('MTM_B200_LIB', 'default'):40s} {tag:10s} {t0.elapsed_time(t1) / K:8.3f} ms / EM iteration",
          flush=True)
    return assign.clone(), params.clone()


a_u, p_u = run("unsorted", z, x, assign0.clone())
key = torch.zeros(n, dtype=torch.int32, device=dev)
for t in range(T):
    key |= torch.isfinite(z[t]).all(dim=1).to(torch.int32) << t
perm = torch.argsort(key, stable=True)
zs, xs = z[:, perm].contiguous(), x[:, perm].contiguous()
a_s, p_s = run("sorted", zs, xs, assign0[perm].contiguous())
back = torch.empty_like(a_s)
back[perm] = a_s
# pattern-grouped but scrambled: aligned groups of 32 keep a uniform pattern, neighbouring groups differ
ng = n // 32
gidx = (torch.arange(ng, device=dev, dtype=torch.int64) * 7919) % ng if ng % 7919 else torch.randperm(ng, device=dev)
perm2 = torch.cat([perm[: ng * 32].view(ng, 32)[gidx].reshape(-1), perm[ng * 32:]])
a_g, p_g = run("grouped", z[:, perm2].contiguous(), x[:, perm2].contiguous(), assign0[perm2].contiguous())
back2 = torch.empty_like(a_g)
back2[perm2] = a_g
print("grouped labels equal:", bool(torch.equal(back2, a_u)))
print("labels equal after 8 iterations:", bool(torch.equal(back, a_u)),
      " max rel param diff:", float(((p_s - p_u).abs() / p_u.abs().clamp_min(1.0)).max()))
