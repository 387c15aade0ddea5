"""Thin object layer over the C ABI: torch tensors carry the device buffers and the stream,
nothing else.  One ``Engine`` per (process, device)."""

from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _capi as C


class DeviceData:
    """A data set resident in HBM: z (T, n, d) or None, x (T, n, l), float64, contiguous.

    ``pattern_order=True`` stores the trajectories in *resident order*: grouped by their per-step "all
    hidden-state coordinates present" pattern (``mtm_pattern_order``), which lets whole warps of the E-step take
    the point-mass shortcut.  ``perm[p]`` is then the caller's index of the trajectory at resident position p;
    every engine call works in resident order, ``to_resident`` / ``from_resident`` translate label vectors and
    (rows, n) result matrices at the boundary.  With ``perm is None`` both orders coincide."""

    perm = None

    def __init__(self, states, observations, device, pattern_order=False):
        def up(a):
            if a is None:
                return None
            if isinstance(a, torch.Tensor):
                t = a.to(device=device, dtype=torch.float64)
            else:
                a = np.ascontiguousarray(np.atleast_3d(np.asarray(a, dtype=np.float64)))
                t = torch.from_numpy(a).to(device)
            return t.contiguous()

        self.z = up(states)
        self.x = up(observations)
        if self.x.dim() != 3 or (self.z is not None and self.z.dim() != 3):
            raise ValueError("states / observations must be (T, n, d|l) arrays")
        self.T, self.n, self.l = (int(v) for v in self.x.shape)
        if self.z is not None:
            if tuple(self.z.shape[:2]) != (self.T, self.n):
                raise ValueError("states and observations disagree on (T, n)")
            self.d = int(self.z.shape[2])
        else:
            self.d = None
        if pattern_order and self.z is not None and self.n > 1:
            self._order(device)

    def _order(self, device):
        eng = Engine.get(device)
        eng._sync_stream()
        perm = torch.empty(self.n, dtype=torch.int64, device=self.z.device)
        st = self.struct()
        eng._chk(eng.lib.mtm_pattern_order(eng.h, ctypes.byref(st), ctypes.c_void_p(perm.data_ptr())))
        for name, w in (("z", self.d), ("x", self.l)):
            src = getattr(self, name)
            dst = torch.empty_like(src)
            eng._chk(eng.lib.mtm_permute_f64(eng.h, ctypes.c_void_p(src.data_ptr()), ctypes.c_void_p(dst.data_ptr()),
                                             self.T, self.n, w, ctypes.c_void_p(perm.data_ptr()), 0))
            setattr(self, name, dst)
            del src  # stream-ordered: the caching allocator reuses the block only after the copy
        self.perm = perm

    def to_resident(self, a, dtype=torch.int32):
        """(n,) or (rows, n) array in the caller's order -> device tensor in resident order."""
        dev = self.x.device
        t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
        t = t.to(device=dev, dtype=dtype).contiguous()
        return t if self.perm is None else self._permute(t, inverse=0)

    def from_resident(self, t):
        """device tensor (n,) or (rows, n) in resident order -> device tensor in the caller's order."""
        if self.perm is None or t is None:
            return t
        return self._permute(t.contiguous(), inverse=1)

    def _permute(self, t, inverse):
        eng = Engine.get(t.device)
        eng._sync_stream()
        rows = 1 if t.dim() == 1 else int(t.shape[0])
        assert int(t.shape[-1]) == self.n and t.dim() <= 2
        out = torch.empty_like(t)
        p = lambda v: ctypes.c_void_p(v.data_ptr())
        if t.dtype == torch.int32:
            eng._chk(eng.lib.mtm_permute_i32(eng.h, p(t), p(out), rows, self.n, p(self.perm), inverse))
        elif t.dtype == torch.float64:
            eng._chk(eng.lib.mtm_permute_f64(eng.h, p(t), p(out), rows, self.n, 1, p(self.perm), inverse))
        else:
            raise TypeError(f"unsupported dtype {t.dtype}")
        return out

    def struct(self, d=None):
        s = C.MtmData()
        s.z_dev = self.z.data_ptr() if self.z is not None else None
        s.x_dev = self.x.data_ptr()
        s.n, s.T, s.l = self.n, self.T, self.l
        s.d = self.d if self.d is not None else int(d)
        return s

    def nbytes(self):
        return (self.z.numel() if self.z is not None else 0) * 8 + self.x.numel() * 8


class Engine:
    _cache: dict = {}
    _sharded: dict = {}
    _local: dict = {}

    @classmethod
    def get_local(cls, device=None) -> "Engine":
        """An engine of `device` that is guaranteed to carry NO communicator: Engine.get() itself unless user code
        attached one to it (dist.attach(Engine.get(...))), else a private second handle.
        Rank-local work of the classes (unsharded models, held-out data every rank holds in full, static regress)
        must never be all-reduced."""
        base = cls.get(device)
        if base.world == 1:
            return base
        if base.device.index not in cls._local:
            cls._local[base.device.index] = Engine(base.device)
        return cls._local[base.device.index]

    @classmethod
    def get_sharded(cls, device=None, group=None) -> "Engine":
        """The engine of `device` whose handle carries the communicator of `group` (created and attached once,
        collectively on all ranks of the group).  Kept apart from Engine.get(): that one never all-reduces, so
        rank-local work (a static regress, held-out data every rank holds in full) stays rank-local."""
        from . import dist as mdist

        base = cls.get(device)
        key = (base.device.index, id(group))
        if key not in cls._sharded:
            eng = Engine(base.device)
            mdist.attach(eng, group)
            cls._sharded[key] = eng
        return cls._sharded[key]

    @classmethod
    def get(cls, device=None) -> "Engine":
        if not torch.cuda.is_available():
            raise C.MtmError("no CUDA device: mtm_b200 runs on a B200 only and has no CPU path")
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        if dev.index not in cls._cache:
            cls._cache[dev.index] = Engine(dev)
        return cls._cache[dev.index]

    def __init__(self, device):
        self.lib = C.load()
        self.device = device
        self.h = ctypes.c_void_p()
        with torch.cuda.device(device):
            self.stream = torch.cuda.current_stream(device)
            C.check(self.lib.mtm_create(device.index, ctypes.c_void_p(self.stream.cuda_stream), ctypes.byref(self.h)))
        self.world, self.rank = 1, 0

    # -- helpers -----------------------------------------------------------------------------
    def _sync_stream(self):
        s = torch.cuda.current_stream(self.device)
        C.check(self.lib.mtm_set_stream(self.h, ctypes.c_void_p(s.cuda_stream)), self.h)

    def _chk(self, rc):
        C.check(rc, self.h)

    def param_stride(self, d, l):
        return int(self.lib.mtm_param_stride(d, l))

    def stats_stride(self, d, l):
        return int(self.lib.mtm_stats_stride(d, l))

    def launch_count(self):
        return int(self.lib.mtm_launch_count(self.h))

    def dev(self, a, dtype):
        if isinstance(a, torch.Tensor):
            return a.to(device=self.device, dtype=dtype).contiguous()
        return torch.from_numpy(np.ascontiguousarray(a)).to(self.device).to(dtype).contiguous()

    # -- operations ----------------------------------------------------------------------------
    def bind(self, data: DeviceData):
        st = data.struct()
        self._chk(self.lib.mtm_set_data(self.h, ctypes.byref(st)))

    def estep(self, data: DeviceData, params, n_c, T0, assign=None, count_switches=False, want_ll=False,
              want_probs=False, want_prenorm=False, d=None, want_info=True):
        """One E-step launch.  Returns dict(assign, ll, probs, prenorm, n_switches, min_count, status)."""
        self._sync_stream()
        st = data.struct(d)
        n = data.n
        params = self.dev(params, torch.float64)
        if assign is None:
            assign = torch.empty(n, dtype=torch.int32, device=self.device)
            count_switches = False
        mk = lambda flag: torch.empty((n_c, n), dtype=torch.float64, device=self.device) if flag else None
        ll, probs, prenorm = mk(want_ll), mk(want_probs), mk(want_prenorm)
        ptr = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
        info = C.EStepInfo()
        self._chk(self.lib.mtm_estep(self.h, ctypes.byref(st), ptr(params), n_c, int(T0), ptr(assign),
                                     1 if count_switches else 0, ptr(ll), ptr(probs), ptr(prenorm),
                                     ctypes.byref(info) if want_info else None))
        return dict(assign=assign, ll=ll, probs=probs, prenorm=prenorm, n_switches=int(info.n_switches),
                    min_count=int(info.min_count), status=int(info.status))

    def estep_prefix(self, data: DeviceData, params, n_c, T0, d=None):
        """(T0, n_c, n) log-likelihoods of every prefix 1..T0 in one launch (mtm_estep_prefix), resident order.
        Returns (tensor, status)."""
        self._sync_stream()
        st = data.struct(d)
        params = self.dev(params, torch.float64)
        T0 = max(1, min(int(T0), data.T))
        out = torch.empty((T0, n_c, data.n), dtype=torch.float64, device=self.device)
        status = ctypes.c_int32(0)
        self._chk(self.lib.mtm_estep_prefix(self.h, ctypes.byref(st), ctypes.c_void_p(params.data_ptr()), n_c, T0,
                                            ctypes.c_void_p(out.data_ptr()), ctypes.byref(status)))
        return out, int(status.value)

    def score(self, data: DeviceData, params, n_c, d=None):
        self._sync_stream()
        st = data.struct(d)
        params = self.dev(params, torch.float64)
        out = (ctypes.c_double * 2)()
        status = ctypes.c_int32(0)
        self._chk(self.lib.mtm_score(self.h, ctypes.byref(st), ctypes.c_void_p(params.data_ptr()), n_c, out,
                                     ctypes.byref(status)))
        return float(out[0]), float(out[1]), int(status.value)

    def mstats(self, data: DeviceData, assign, n_c):
        self._sync_stream()
        self.bind(data)
        stats = torch.empty((n_c, self.stats_stride(data.d, data.l)), dtype=torch.float64, device=self.device)
        self._chk(self.lib.mtm_mstats(self.h, ctypes.c_void_p(assign.data_ptr()), n_c, ctypes.c_void_p(stats.data_ptr())))
        return stats

    def finalize(self, stats, n_c, d, l, n_total, alpha):
        self._sync_stream()
        params = torch.empty((n_c, self.param_stride(d, l)), dtype=torch.float64, device=self.device)
        status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._chk(self.lib.mtm_finalize(self.h, ctypes.c_void_p(stats.data_ptr()), n_c, d, l, int(n_total), float(alpha),
                                        ctypes.c_void_p(params.data_ptr()), ctypes.c_void_p(status.data_ptr())))
        return params, status

    def mstep(self, data: DeviceData, assign, n_c, alpha):
        self._sync_stream()
        self.bind(data)
        params = torch.empty((n_c, self.param_stride(data.d, data.l)), dtype=torch.float64, device=self.device)
        info = C.EStepInfo()
        self._chk(self.lib.mtm_mstep(self.h, ctypes.c_void_p(assign.data_ptr()), n_c, float(alpha),
                                     ctypes.c_void_p(params.data_ptr()), ctypes.byref(info)))
        return params, int(info.min_count), int(info.status)

    def em_fit(self, data: DeviceData, assign, n_c, alpha, n_steps, params=None):
        """train(): assign (int32 device tensor) is updated in place.  Returns (params, info dict)."""
        self._sync_stream()
        self.bind(data)
        if params is None:
            params = torch.full((n_c, self.param_stride(data.d, data.l)), float("nan"), dtype=torch.float64,
                                device=self.device)
        info = C.FitInfo()
        cap = max(1, int(n_steps))
        hist = (ctypes.c_int64 * cap)()
        self._chk(self.lib.mtm_em_fit(self.h, n_c, float(alpha), int(n_steps), ctypes.c_void_p(assign.data_ptr()),
                                      ctypes.c_void_p(params.data_ptr()), ctypes.byref(info), hist, cap))
        return params, dict(exit_code=int(info.exit_code), exit=C.EXIT_NAMES.get(int(info.exit_code), "?"),
                            n_e_steps=int(info.n_e_steps), n_m_steps=int(info.n_m_steps), status=int(info.status),
                            last_switches=int(info.last_switches), min_count=int(info.min_count),
                            switches=[int(hist[k]) for k in range(min(cap, int(info.n_e_steps)))])

    def mstep_ext(self, data: DeviceData, assign, n_c, alpha):
        """Component-style fits (framework_extended's StateSpaceLinearGaussian.fit) of every label value in one pass
        (mtm_mstep_ext).  Returns (params (n_c, stride), min_count, status)."""
        self._sync_stream()
        self.bind(data)
        params = torch.empty((n_c, self.param_stride(data.d, data.l)), dtype=torch.float64, device=self.device)
        info = C.EStepInfo()
        self._chk(self.lib.mtm_mstep_ext(self.h, ctypes.c_void_p(assign.data_ptr()), n_c, float(alpha),
                                         ctypes.c_void_p(params.data_ptr()), ctypes.byref(info)))
        return params, int(info.min_count), int(info.status)

    def em_fit_ext(self, data: DeviceData, assign, n_c, alpha, n_steps):
        """The EM loop of framework_extended's StateSpaceMixtureModel.fit with linear-Gaussian components
        (mtm_em_fit_ext): assign (int32 device tensor) is updated in place.  Returns (params, info dict)."""
        self._sync_stream()
        self.bind(data)
        params = torch.full((n_c, self.param_stride(data.d, data.l)), float("nan"), dtype=torch.float64,
                            device=self.device)
        prev = torch.empty_like(assign)
        info = C.FitInfo()
        cap = max(1, int(n_steps))
        hist = (ctypes.c_int64 * cap)()
        self._chk(self.lib.mtm_em_fit_ext(self.h, n_c, float(alpha), int(n_steps), ctypes.c_void_p(assign.data_ptr()),
                                          ctypes.c_void_p(prev.data_ptr()), ctypes.c_void_p(params.data_ptr()),
                                          ctypes.byref(info), hist, cap))
        return params, dict(exit_code=int(info.exit_code), exit=C.EXIT_NAMES.get(int(info.exit_code), "?"),
                            n_e_steps=int(info.n_e_steps), n_m_steps=int(info.n_m_steps), status=int(info.status),
                            last_switches=int(info.last_switches), min_count=int(info.min_count),
                            switches=[int(hist[k]) for k in range(min(cap, int(info.n_e_steps)))])

    def em_step(self, data: DeviceData, assign, params, n_c, alpha, want_info=True):
        """One EM iteration on resident data (the reference's E_step(); M_step() pair): assign and params
        (device tensors) are updated in place.  Returns (n_switches, min_count, status) if want_info."""
        self._sync_stream()
        self.bind(data)
        info = C.EStepInfo()
        self._chk(self.lib.mtm_em_step(self.h, n_c, float(alpha), ctypes.c_void_p(assign.data_ptr()),
                                       ctypes.c_void_p(params.data_ptr()), 3, ctypes.byref(info) if want_info else None))
        return (int(info.n_switches), int(info.min_count), int(info.status)) if want_info else None

    def em_step_host(self, z_host, x_host, z_dev, x_dev, assign_host, assign_dev, params_dev, params_host, n_c, alpha,
                     n_slabs=16):
        """One EM iteration end to end from (pinned) host tensors: slab-pipelined upload, fused E-step +
        statistics per slab, labels streamed back, parameters solved and downloaded (mtm_em_step_host)."""
        self._sync_stream()
        T, n, d = (int(v) for v in z_host.shape)
        l = int(x_host.shape[2])
        info = C.EStepInfo()
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        self._chk(self.lib.mtm_em_step_host(self.h, p(z_host), p(x_host), T, n, d, l, p(z_dev), p(x_dev), n_c,
                                            float(alpha), p(assign_host), p(assign_dev), p(params_dev), p(params_host),
                                            int(n_slabs), ctypes.byref(info)))
        return int(info.n_switches), int(info.min_count), int(info.status)

    def sweep_fit(self, data: DeviceData, assign, n_c, alpha, n_steps, want_scores=True):
        """Batched restarts (mtm_sweep_fit): assign is an (R, n) int32 device tensor, updated in place.
        Returns (params (R, n_c, stride), infos [R dicts], scores (R, 2) ndarray or None, score_status)."""
        self._sync_stream()
        self.bind(data)
        R = int(assign.shape[0])
        params = torch.full((R, n_c, self.param_stride(data.d, data.l)), float("nan"), dtype=torch.float64,
                            device=self.device)
        infos = (C.FitInfo * R)()
        scores = (ctypes.c_double * (2 * R))() if want_scores else None
        sstat = (ctypes.c_int32 * R)() if want_scores else None
        cap = max(1, int(n_steps))
        hist = (ctypes.c_int64 * (cap * R))()
        self._chk(self.lib.mtm_sweep_fit(self.h, n_c, R, float(alpha), int(n_steps), ctypes.c_void_p(assign.data_ptr()),
                                         ctypes.c_void_p(params.data_ptr()), infos, scores, sstat, hist, cap))
        out = []
        for r in range(R):
            i = infos[r]
            out.append(dict(exit_code=int(i.exit_code), exit=C.EXIT_NAMES.get(int(i.exit_code), "?"),
                            n_e_steps=int(i.n_e_steps), n_m_steps=int(i.n_m_steps), status=int(i.status),
                            last_switches=int(i.last_switches), min_count=int(i.min_count),
                            switches=[int(hist[r * cap + k]) for k in range(min(cap, int(i.n_e_steps)))]))
        sc = np.array(scores[:], dtype=np.float64).reshape(R, 2) if want_scores else None
        ss = np.array(sstat[:], dtype=np.int32) if want_scores else None
        return params, out, sc, ss

    def random_labels(self, seed, n_c, n, out=None):
        """default_rng(seed).integers(0, n_c, n) generated on the device, bit-identical to numpy
        (mtm_random_labels).  Returns (int32 device tensor, number of rejected draws replayed)."""
        self._sync_stream()
        st = np.random.PCG64(seed).state["state"]
        m = (1 << 64) - 1
        arr = (ctypes.c_uint64 * 4)(st["state"] >> 64, st["state"] & m, st["inc"] >> 64, st["inc"] & m)
        if out is None:
            out = torch.empty(int(n), dtype=torch.int32, device=self.device)
        rej = ctypes.c_int(0)
        self._chk(self.lib.mtm_random_labels(self.h, arr, int(n_c), int(n), ctypes.c_void_p(out.data_ptr()),
                                             ctypes.byref(rej)))
        return out, int(rej.value)

    def kmeans_init(self, data: DeviceData, n_c, seed, max_iter=300, rel_tol=1e-4):
        """The reference's init="kmeans" (MMM:132-151) on the device (mtm_kmeans_init): k-means of the first
        non-null states, sklearn's algorithm on sklearn's random numbers — they are drawn here from numpy's
        RandomState(seed) in the order sklearn draws them.  Returns (int32 labels in the order of `data`, info)."""
        self._sync_stream()
        n = data.n
        rs = np.random.RandomState(seed)
        first = int(rs.choice(n, p=np.full(n, 1.0) / float(n)))      # _kmeans_plusplus: first centre
        n_trials = 2 + int(np.log(n_c))
        u = np.ascontiguousarray(np.stack([rs.uniform(size=n_trials) for _ in range(1, n_c)]).ravel()
                                 if n_c > 1 else np.zeros(1))
        out = torch.empty(n, dtype=torch.int32, device=self.device)
        info = C.KMeansInfo()
        st = data.struct()
        perm = ctypes.c_void_p(data.perm.data_ptr()) if data.perm is not None else None
        self._chk(self.lib.mtm_kmeans_init(self.h, ctypes.byref(st), perm, int(n_c), first,
                                           u.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), n_trials, int(max_iter),
                                           float(rel_tol), ctypes.c_void_p(out.data_ptr()), ctypes.byref(info)))
        return out, dict(n_iter=int(info.n_iter), converged=int(info.converged), empty_cluster=int(info.empty_cluster),
                         tol=float(info.tol))

    def fingerprint(self, data: DeviceData, d=None) -> str:
        """32 hex digits: the device-side key of a data set in the caller's order (mtm_fingerprint)."""
        self._sync_stream()
        st = data.struct(d if d is not None else (data.d or 1))
        key = (ctypes.c_uint64 * 2)()
        perm = ctypes.c_void_p(data.perm.data_ptr()) if data.perm is not None else None
        self._chk(self.lib.mtm_fingerprint(self.h, ctypes.byref(st), perm, key))
        return f"{key[0]:016x}{key[1]:016x}"

    def set_delta_stats(self, mode=1):
        """Delta statistics of the fused EM iteration (mtm_set_delta_stats): 0 / False off, 1 / True automatic
        (the default: on unless the set is launch-latency bound), 2 on for every size."""
        self._chk(self.lib.mtm_set_delta_stats(self.h, int(mode)))

    def release_workspace(self):
        """Give back the grow-only workspace of pattern_order / kmeans_init (mtm_release_workspace)."""
        self._chk(self.lib.mtm_release_workspace(self.h))

    def fp64_peak(self, reps=10):
        self._sync_stream()
        out = ctypes.c_double(0.0)
        self._chk(self.lib.mtm_fp64_peak(self.h, reps, ctypes.byref(out)))
        return float(out.value)

    # -- sharding ------------------------------------------------------------------------------
    def unique_id(self) -> bytes:
        buf = ctypes.create_string_buffer(128)
        C.check(self.lib.mtm_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, id_bytes: bytes, world: int, rank: int):
        buf = ctypes.create_string_buffer(id_bytes, 128)
        self._chk(self.lib.mtm_comm_init(self.h, buf, world, rank))
        self.world, self.rank = world, rank

    def comm_attach(self, nccl_comm: int, world: int, rank: int):
        """Use a communicator the caller owns (an ncclComm_t as an integer address); see mtm_comm_attach."""
        self._chk(self.lib.mtm_comm_attach(self.h, ctypes.c_void_p(nccl_comm), world, rank))
        self.world, self.rank = world, rank

    def p2p_export(self) -> bytes:
        buf = ctypes.create_string_buffer(64)
        self._chk(self.lib.mtm_p2p_export(self.h, buf))
        return buf.raw

    def p2p_attach(self, handles, world: int, rank: int):
        blob = b"".join(handles)
        buf = ctypes.create_string_buffer(blob, len(blob))
        self._chk(self.lib.mtm_p2p_attach(self.h, buf, world, rank))

    def comm_destroy(self):
        self._chk(self.lib.mtm_comm_destroy(self.h))
        self.world, self.rank = 1, 0
