"""Drop-in for the reference's ``framework/marginalizable_mixture_model.py`` (MMM): the
hard-assignment EM of a mixture of linear-Gaussian state-space models, with every number
computed by the sm_100a kernels behind ``include/mtm_b200.h``.

Same class name, constructor, attributes, method names, argument meaning and error behaviour as
the reference class (MMM:39) for the EM / predict path; reporting and plotting helpers of the
reference (MMM:321-437, 1473-1916) are presentation code and are not part of this path.

There is no CPU fallback: without the CUDA library / a B200 every numeric method raises.
"""

from __future__ import annotations

import datetime
import glob
import gzip
import hashlib
import os
import pickle
import string

import numpy as np
import torch

try:
    from .. import _capi as C
    from .. import dist as mdist
    from ..engine import DeviceData, Engine
except ImportError:
    # Imported as the TOP-LEVEL package `framework` (the package directory was put first on sys.path so that
    # the reference's `from framework import marginalizable_mixture_model as mixmodel` resolves here,
    # INTEGRATION.md section 1): reach the rest of the package through its importable alias.
    import sys as _sys

    _root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if _root not in _sys.path:
        _sys.path.insert(0, _root)
    import mtm_b200 as _pkg  # noqa: F401
    from mtm_b200 import _capi as C
    from mtm_b200 import dist as mdist
    from mtm_b200.engine import DeviceData, Engine

np_eps = np.finfo(float).eps
# reference: <repo>/tmp (MMM:36, 229).  Configurable because a package directory may be read-only.
home_dir = os.environ.get("MTM_B200_HOME", os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _raise_for_status(status: int, alpha: float = 0.0):
    """Re-raise what the reference raises for the same numerical condition."""
    if status & C.STATUS_COMM_TIMEOUT:
        raise C.MtmError("peer-memory all-reduce timed out: a rank did not arrive")
    if status & C.STATUS_NONFINITE:
        # numba's det/solve on a covariance block holding NaN/inf (SSM:771-773)
        raise np.linalg.LinAlgError("Array must not contain infs or NaNs.")
    if status & C.STATUS_SINGULAR:
        raise np.linalg.LinAlgError("Matrix is singular to machine precision.")
    if status & C.STATUS_EMPTY_FIT:
        if alpha > 2 * np_eps:  # sklearn Ridge.fit on 0 samples (MMM:1226-1227)
            raise ValueError("Found array with 0 sample(s) while a minimum of 1 is required by Ridge.")
        raise np.linalg.LinAlgError("Arrays cannot be empty")  # numba lstsq (MMM:1194)
    if status & C.STATUS_CHOL_FAIL:
        raise np.linalg.LinAlgError("rank-deficient regression design: normal equations not positive definite")


class MMLinGaussSS_marginalizable:
    """Probabilistic mixture of linear-Gaussian state-space models trained by hard-assignment EM
    (reference: MMM:39-1471)."""

    def __init__(
        self,
        n_clusters: int,
        states: np.array,
        observations: np.array,
        random_seed: int = 42,
        init: str = "random",
        alpha: float = 0.0,
        *,
        device=None,
        group=None,
        _shared_data: DeviceData = None,
    ):
        """Same arguments as the reference constructor (MMM:45-53); ``device`` (keyword-only,
        optional) picks the CUDA device.

        Multi-GPU: when a ``torch.distributed`` process group is initialised (one process per GPU, e.g. under
        torchrun) every rank constructs the model with the SAME full arrays, exactly like the single-process
        script, and the class shards the trajectory axis itself: a rank keeps rows
        ``dist.shard_bounds(n, world, rank)`` on its GPU, every fit / score is one collective call with one
        all-reduce of the sufficient statistics per EM iteration, and labels / (n_c, n) matrices are all-gathered,
        so that every rank ends with the reference's full ``cluster_assignment`` and identical parameters.
        ``group``: the process group to shard over (default: the world group); ``group=False`` keeps the model
        local to this process."""
        states, observations = map(np.atleast_3d, (states, observations))
        self.n_clusters = int(n_clusters)
        self.states = np.array(states)
        self.observations = np.array(observations)
        self.n_timesteps, self.n_data, self.d_states = self.states.shape
        self.d_observations = self.observations.shape[-1]
        d, l, k = self.d_states, self.d_observations, self.n_clusters

        # Placeholder parameters from the unseeded global RNG, as the reference draws them
        # (MMM:83-125); `train` overwrites them in its first M-step.
        self.cluster_propensities = np.ones(shape=[k]) / k
        self.init_state_means = [np.random.normal(size=[d]) for _ in range(k)]
        spd = lambda m: m @ m.T + np.eye(m.shape[0])
        self.init_state_covs = [spd(np.random.normal(size=[d, d])) for _ in range(k)]
        self.transition_matrices = [np.random.normal(size=[d, d]) for _ in range(k)]
        self.transition_covs = [spd(np.random.normal(size=[d, d])) for _ in range(k)]
        self.measurement_matrices = [np.random.normal(size=[d, l]) for _ in range(k)]
        self.measurement_covs = [spd(np.random.normal(size=[l, l])) for _ in range(k)]

        self.random_seed = random_seed
        self.rng = np.random.default_rng(seed=self.random_seed)
        self.init = init
        self.alpha = alpha if alpha > 2 * np_eps else 0

        self._correspondence = dict(zip(range(k), string.ascii_uppercase))
        self.inverse_correspondence = {v: c for c, v in self._correspondence.items()}
        self._hex_hash = None  # md5 over all data bytes (MMM:177-188): computed on first use, see hex_hash
        self.time_stamp = (datetime.datetime.now(datetime.timezone.utc).replace(microsecond=0)
                           .astimezone().isoformat())
        self.last_trained = None

        self._group = group
        self._world, self._rank = mdist.active_group(group)
        if self._world > 1 and self.n_data < 32 * self._world:
            self._world, self._rank = 1, 0  # too small to shard: every rank fits its own full copy
        if self._world > 1 and device is None and "LOCAL_RANK" in os.environ:
            device = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
        self._device = device
        self._data = _shared_data  # uploaded lazily
        self.fit_info = None       # dict from the last train(): exit, n_e_steps, switches, ...
        self.kmeans_info = None    # dict from a k-means start computed on the device
        self.cluster_assignment = self._initial_assignment()

    # ------------------------------------------------------------------ initial assignment
    def _initial_assignment(self):
        """Host-side, exactly the reference's three initialisers (MMM:131-168)."""
        import sklearn.cluster as skl_cluster

        match self.init:
            case "k-means" | "kmeans" if self._kmeans_on_device():
                # large data: the same k-means (sklearn's algorithm on sklearn's random numbers, mtm_kmeans_init)
                # on the resident set instead of a host KMeans over n rows
                data = self._train_data()
                labels, self.kmeans_info = self.engine.kmeans_init(data, self.n_clusters, self.random_seed)
                return data.from_resident(labels).cpu().numpy().astype(np.int32)
            case "k-means" | "kmeans":
                idx = np.argmax(np.isfinite(self.states).all(axis=2), axis=0).ravel()
                first = self.states[idx, np.arange(self.n_data), :]
                with np.errstate(all="ignore"):
                    first = np.where(np.isfinite(first), first, np.nanmean(first, axis=0, keepdims=True))
                return skl_cluster.KMeans(n_clusters=self.n_clusters, init="k-means++",
                                          random_state=self.random_seed).fit_predict(first)
            case "kmeans-all" | "k-means-all":
                flat = np.transpose(self.states, (1, 0, 2)).reshape(self.n_data, -1)
                return skl_cluster.KMeans(n_clusters=self.n_clusters, init="k-means++",
                                          random_state=self.random_seed).fit_predict(flat)
            case _:
                return self.rng.integers(low=0, high=self.n_clusters, size=self.n_data)

    def _kmeans_on_device(self) -> bool:
        """k-means starts of large, unsharded sets run on the device; MTM_B200_KMEANS=device|host forces either."""
        mode = os.environ.get("MTM_B200_KMEANS", "auto").lower()
        if self._world > 1 or mode == "host" or not isinstance(self.random_seed, (int, np.integer)):
            return False
        return mode == "device" or self.n_data >= 200_000

    # ------------------------------------------------------------------ small properties
    @property
    def n_free_params(self) -> int:
        """MMM:197-214."""
        tri = lambda m: len(np.triu_indices_from(np.atleast_2d(m))[0])
        return (sum(np.size(x) for x in [self.cluster_propensities] + list(self.init_state_means)
                    + list(self.transition_matrices) + list(self.measurement_matrices))
                + sum(tri(x) for x in list(self.init_state_covs) + list(self.transition_covs)
                      + list(self.measurement_covs)))

    @property
    def hex_hash(self) -> str:
        """Cache key of (data, n_clusters, alpha), evaluated lazily — only the pickle cache needs it.

        Small data: the reference's key, md5 of states bytes + observations bytes + n_clusters (+ alpha)
        (MMM:177-188), so caches are interchangeable with the reference's.  Data of 64 MB and more (and not sharded):
        the md5 is taken over the 128-bit device fingerprint of the resident data set (``mtm_fingerprint``, one pass
        at HBM speed) instead of over 5 GB of host bytes; ``reference_hex_hash`` still gives the reference's key on
        demand, ``MTM_B200_HASH=md5`` makes it the default again."""
        if self._hex_hash is None:
            big = (self.states.size + self.observations.size) * 8 >= (64 << 20)
            if big and self._world == 1 and os.environ.get("MTM_B200_HASH", "fast").lower() != "md5":
                md5 = hashlib.md5()
                md5.update(b"mtm_b200-fingerprint:" + self.engine.fingerprint(self._train_data()).encode("ascii"))
                self._hex_hash = self._hash_suffix(md5).hexdigest()
            else:
                self._hex_hash = self.reference_hex_hash
        return self._hex_hash

    def _hash_suffix(self, md5):
        md5.update(str(self.n_clusters).encode("utf-8"))
        if self.alpha > 2 * np_eps:
            md5.update(np.format_float_positional(self.alpha, unique=True).encode("utf-8"))
        return md5

    @property
    def reference_hex_hash(self) -> str:
        """md5 of states bytes + observations bytes + n_clusters (+ alpha): the reference's key (MMM:177-188)."""
        md5 = hashlib.md5()
        md5.update(np.ascontiguousarray(self.states).data)
        md5.update(np.ascontiguousarray(self.observations).data)
        return self._hash_suffix(md5).hexdigest()

    @hex_hash.setter
    def hex_hash(self, value: str) -> None:
        self._hex_hash = value

    @property
    def correspondence(self) -> dict:
        return self._correspondence

    @correspondence.setter
    def correspondence(self, corr: dict) -> None:
        self._correspondence = corr
        self.inverse_correspondence = {v: k for k, v in self._correspondence.items()}

    # ------------------------------------------------------------------ device plumbing
    @property
    def engine(self) -> Engine:
        """The engine of the training set: with the group's communicator when the model is sharded."""
        if self._world > 1:
            return Engine.get_sharded(self._device, self._group if self._group is not False else None)
        return Engine.get_local(self._device)

    def _engine_for(self, data) -> Engine:
        return self.engine if getattr(data, "shard", None) else Engine.get_local(self._device)

    def _shard_of(self, n):
        """(lo, hi) of this rank for an n-trajectory set, or None if the set is not sharded."""
        if self._world == 1 or n < 32 * self._world:
            return None
        return mdist.shard_bounds(n, self._world, self._rank)

    def _make_data(self, states, observations, pattern_order=False) -> DeviceData:
        """Device copy of (this rank's rows of) a data set; ``data.shard`` = (n, lo, hi) when sharded."""
        n = observations.shape[1]
        sh = self._shard_of(n)
        if sh is not None:
            states = None if states is None else np.ascontiguousarray(states[:, sh[0]:sh[1]])
            observations = np.ascontiguousarray(observations[:, sh[0]:sh[1]])
        data = DeviceData(states, observations, Engine.get(self._device).device, pattern_order=pattern_order)
        data.shard = None if sh is None else (n, sh[0], sh[1])
        return data

    def _gather(self, data, t):
        """(..., n_rank) device tensor in the caller's order -> (..., n) on every rank."""
        sh = getattr(data, "shard", None)
        if sh is None or t is None:
            return t
        return mdist.gather_last(t, sh[0], self._world, self._rank, self._group if self._group is not False else None)

    def _train_data(self) -> DeviceData:
        if self._data is None:
            # resident order = grouped by the per-step missing-coordinate pattern (engine.DeviceData); everything
            # that leaves this class is translated back to the caller's order
            self._data = self._make_data(self.states, self.observations, pattern_order=True)
        return self._data

    def _packed_params(self) -> np.ndarray:
        """(n_c, 1+d+3d^2+dl+l^2) rows [pi | m | S | A | Gamma | H | Lambda] (C-ABI layout)."""
        d, l = self.d_states, self.d_observations
        rows = []
        for c in range(self.n_clusters):
            rows.append(np.concatenate([
                [self.cluster_propensities[c]],
                np.reshape(self.init_state_means[c], d),
                np.reshape(self.init_state_covs[c], d * d),
                np.reshape(self.transition_matrices[c], d * d),
                np.reshape(self.transition_covs[c], d * d),
                np.reshape(self.measurement_matrices[c], d * l),
                np.reshape(self.measurement_covs[c], l * l)]))
        return np.ascontiguousarray(np.stack(rows), dtype=np.float64)

    def _unpack_params(self, packed) -> None:
        P = packed.detach().cpu().numpy() if isinstance(packed, torch.Tensor) else np.asarray(packed)
        d, l = self.d_states, self.d_observations
        o = [0]

        def take(shape):
            size = int(np.prod(shape))
            out = [P[c, o[0]:o[0] + size].reshape(shape).copy() for c in range(self.n_clusters)]
            o[0] += size
            return out

        self.cluster_propensities = P[:, 0].copy()
        o[0] = 1
        self.init_state_means = take((d,))
        self.init_state_covs = take((d, d))
        self.transition_matrices = take((d, d))
        self.transition_covs = take((d, d))
        self.measurement_matrices = take((d, l))
        self.measurement_covs = take((l, l))

    def _assign_dev(self):
        """Current labels (this rank's rows) as an int32 device tensor in the training set's resident order."""
        data = self._train_data()
        lab = self.cluster_assignment
        if getattr(data, "shard", None):
            lab = lab[data.shard[1]:data.shard[2]]
        return data.to_resident(np.ascontiguousarray(lab, dtype=np.int32))

    def _assign_host(self, assign_dev) -> np.ndarray:
        data = self._train_data()
        return self._gather(data, data.from_resident(assign_dev)).cpu().numpy().astype(np.int64)

    def _resolve(self, states, observations):
        """The reference's data-override rule: if `states` is None BOTH come from the training
        set (MMM:469-471, 756-758); otherwise the given arrays are used."""
        if states is None:
            return self._train_data()
        if observations is None:
            raise TypeError("observations must be given together with states")
        return self._make_data(np.atleast_3d(np.asarray(states, dtype=np.float64)),
                               np.atleast_3d(np.asarray(observations, dtype=np.float64)))

    # ------------------------------------------------------------------ likelihoods
    def _estep(self, data, T0, *, ll=False, probs=False, prenorm=False, d=None):
        res = self._engine_for(data).estep(data, self._packed_params(), self.n_clusters, T0, want_ll=ll,
                                           want_probs=probs, want_prenorm=prenorm, d=d)
        status = res["status"]
        if getattr(data, "shard", None):  # every rank must raise (or not) together
            status = self._or_status(status)
        _raise_for_status(status, self.alpha)
        for k in ("assign", "ll", "probs", "prenorm"):  # resident order -> caller's order (-> all ranks' rows)
            res[k] = self._gather(data, data.from_resident(res[k]))
        return res

    def _or_status(self, status: int) -> int:
        """Union of the ranks' status bits (each bit summed separately)."""
        grp = self._group if self._group is not False else None
        dev = Engine.get(self._device).device
        return sum(b for b in (1, 2, 4, 8, 16) if mdist.sum_int(1 if status & b else 0, dev, grp) > 0)

    def conditional_log_likelihoods_first_T0_steps(self, c: int, T0: int, *, states=None, observations=None):
        """log p(z_i, x_i | cluster c) on steps 1..T0 (MMM:439-494)."""
        assert 1 <= T0 <= self.n_timesteps
        data = self._resolve(states, observations)
        return self._estep(data, T0, ll=True)["ll"][c].cpu().numpy()

    def conditional_log_likelihoods(self, c: int, *, states=None, observations=None):
        """MMM:496-530."""
        return self.conditional_log_likelihoods_first_T0_steps(c, self.n_timesteps, states=states,
                                                               observations=observations)

    def log_likelihood_matrix(self, *, states=None, observations=None, T0=None) -> np.ndarray:
        """(n_clusters, n) matrix of conditional log-likelihoods in ONE launch (the reference needs
        n_clusters calls of `conditional_log_likelihoods`)."""
        data = self._resolve(states, observations)
        return self._estep(data, self.n_timesteps if T0 is None else T0, ll=True)["ll"].cpu().numpy()

    def mle_cluster_assignment(self, *, return_probs: bool = False, return_prenormalized_log_probs: bool = False,
                               states=None, observations=None):
        """argmax_c pi_c exp(ll_c) [, probs (n_c, n)] [, log pi_c + ll_c (n_c, n)]  (MMM:724-791)."""
        data = self._resolve(states, observations)
        want_p = return_probs or return_prenormalized_log_probs
        res = self._estep(data, self.n_timesteps, probs=want_p, prenorm=return_prenormalized_log_probs)
        assignments = res["assign"].cpu().numpy().astype(np.int64)
        if not want_p:
            return assignments
        probs = res["probs"].cpu().numpy()
        if not return_prenormalized_log_probs:
            return assignments, probs
        return assignments, probs, res["prenorm"].cpu().numpy()

    def cluster_assignment_index(self, *, cluster: str = "A", states=None, observations=None):
        """Pre-normalised log-odds of assignment to `cluster` (MMM:793-806)."""
        return self.mle_cluster_assignment(states=states, observations=observations, return_probs=True,
                                           return_prenormalized_log_probs=True)[-1][self.inverse_correspondence[cluster]]

    def cluster_propensities_over_time(self, *, states=None, observations=None):
        """(T, n, n_c) membership probabilities from the first t+1 steps (MMM:532-577)."""
        data = self._resolve(states, observations)
        return self._propensities_over_time(data, min(self.n_timesteps, data.T), None)

    def _propensities_over_time(self, data, T0, d):
        """All T0 prefixes from ONE launch (mtm_estep_prefix): the sequential filter passes through every
        prefix likelihood anyway; the reference re-evaluates each prefix from scratch."""
        L, status = self._engine_for(data).estep_prefix(data, self._packed_params(), self.n_clusters, T0, d=d)
        if getattr(data, "shard", None):
            status = self._or_status(status)
        _raise_for_status(status, self.alpha)
        L = self._gather(data, data.from_resident(L.reshape(T0 * self.n_clusters, data.n)))
        L = L.reshape(T0, self.n_clusters, -1)
        pc_t = self.cluster_propensities[None, None, :] * np.exp(L.cpu().numpy().transpose(0, 2, 1))
        pc_t /= np.sum(pc_t, axis=-1, keepdims=True)
        assert np.all(pc_t >= 0.0) and np.allclose(np.sum(pc_t, axis=-1), 1.0)
        return pc_t

    def initial_full_data_cluster_assignment(self, *, states=None, observations=None):
        """Assignment from the first time step only (MMM:909-941)."""
        data = self._resolve(states, observations)
        return self._estep(data, 1)["assign"].cpu().numpy().astype(np.int64)

    # ---- per-visit snapshots without materialising them -------------------------------------------
    def _snapshot_rows(self, s: int) -> np.ndarray:
        """Parameter rows whose initial state is the model's marginal at time step s: (m A^s, V_s) with
        V_0 = S, V_t = A' V_{t-1} A + Gamma (SSM:62-64, 312-314).  A one-step evaluation of x_s, z_s under such a
        row is the likelihood of the trajectory whose every other step is missing."""
        P = self._packed_params()
        d = self.d_states
        for c in range(self.n_clusters):
            m, S = np.array(self.init_state_means[c], dtype=float).reshape(d), np.array(self.init_state_covs[c], dtype=float).reshape(d, d)
            A, G = np.reshape(self.transition_matrices[c], (d, d)), np.reshape(self.transition_covs[c], (d, d))
            for _ in range(s):
                m, S = m @ A, A.T @ S @ A + G
            P[c, 1:1 + d], P[c, 1 + d:1 + d + d * d] = m, S.ravel()
        return P

    def mle_cluster_assignment_snapshots(self, *, states, observations, return_probs: bool = False,
                                         return_prenormalized_log_probs: bool = False, hidden_states: bool = True):
        """``mle_cluster_assignment`` on the per-visit snapshots of (states, observations) — the n T pseudo-trajectories
        ``util.parcellate_arrays`` builds (util/util_state_space.py:417-440; inference-adni-xval.py:194-221), each
        with ONE time step present — without building them: snapshot s of all trajectories is one single-step
        E-step launch on the views ``states[s]``, ``observations[s]`` under parameter rows whose initial state is
        the model's marginal at step s.  T launches over n single steps instead of one launch over n T trajectories
        of T steps.  Output layout = the parcellated call's: position s n + i is snapshot s of trajectory i.
        ``hidden_states=False`` marginalises the hidden states (the call with ``np.nan * np.ones_like(z_parc)``)."""
        states = np.atleast_3d(np.asarray(states, dtype=np.float64))
        observations = np.atleast_3d(np.asarray(observations, dtype=np.float64))
        eng = Engine.get_local(self._device)
        full = DeviceData(states if hidden_states else None, observations, eng.device)
        want_p = return_probs or return_prenormalized_log_probs
        labels, probs, prenorm = [], [], []
        for s in range(full.T):
            view = DeviceData.__new__(DeviceData)
            view.z = full.z[s:s + 1] if hidden_states else None
            view.x, view.T, view.n, view.l, view.perm = full.x[s:s + 1], 1, full.n, full.l, None
            view.d = full.d if hidden_states else None
            res = eng.estep(view, self._snapshot_rows(s), self.n_clusters, 1, want_probs=want_p,
                            want_prenorm=return_prenormalized_log_probs, d=self.d_states)
            _raise_for_status(res["status"], self.alpha)
            labels.append(res["assign"])
            probs.append(res["probs"])
            prenorm.append(res["prenorm"])
        a = torch.cat(labels).cpu().numpy().astype(np.int64)
        if not want_p:
            return a
        pr = torch.cat(probs, dim=1).cpu().numpy()
        if not return_prenormalized_log_probs:
            return a, pr
        return a, pr, torch.cat(prenorm, dim=1).cpu().numpy()

    def cluster_assignment_index_snapshots(self, *, cluster: str = "A", states, observations, hidden_states: bool = True):
        """``cluster_assignment_index`` on the per-visit snapshots (see ``mle_cluster_assignment_snapshots``)."""
        return self.mle_cluster_assignment_snapshots(
            states=states, observations=observations, return_probs=True, return_prenormalized_log_probs=True,
            hidden_states=hidden_states)[-1][self.inverse_correspondence[cluster]]

    # ---- observation-only route (hidden states marginalised) ------------------------------
    def _obs_data(self, observations):
        if observations is None:
            tr = self._train_data()
            od = DeviceData.__new__(DeviceData)
            od.z, od.x, od.T, od.n, od.l, od.d, od.perm = None, tr.x, tr.T, tr.n, tr.l, None, tr.perm
            od.shard = getattr(tr, "shard", None)
            return od
        return self._make_data(None, np.atleast_3d(np.asarray(observations, dtype=np.float64)))

    def observed_condl_log_lik_first_T0_steps(self, c: int, T0: int, *, observations=None):
        """log p(x_i | cluster c) on steps 1..T0 (MMM:997-1048)."""
        assert 1 <= T0 <= self.n_timesteps
        return self._estep(self._obs_data(observations), T0, ll=True, d=self.d_states)["ll"][c].cpu().numpy()

    def observed_conditional_log_likelihoods(self, c: int, observations=None):
        """MMM:1050-1075."""
        return self.observed_condl_log_lik_first_T0_steps(c, self.n_timesteps, observations=observations)

    def observed_cluster_propensities_over_time(self, observations=None):
        """MMM:1077-1123."""
        data = self._obs_data(observations)
        return self._propensities_over_time(data, data.T, self.d_states)

    def observations_mle_cluster_assignment(self, *, return_probs: bool = False, observations=None):
        """argmax over p(x | c) pi_c  [, probs]  (MMM:1125-1165)."""
        res = self._estep(self._obs_data(observations), self.n_timesteps, probs=return_probs, d=self.d_states)
        a = res["assign"].cpu().numpy().astype(np.int64)
        return (a, res["probs"].cpu().numpy()) if return_probs else a

    # ---- scores --------------------------------------------------------------------------------
    def _score(self, states, observations):
        data = self._resolve(states, observations)
        # sharded data: the library all-reduces the two sums; data every rank holds in full: the local engine
        q, ll, status = self._engine_for(data).score(data, self._packed_params(), self.n_clusters)
        if getattr(data, "shard", None):
            status = self._or_status(status)
        _raise_for_status(status, self.alpha)
        return q, ll

    def e_complete_data_log_lik(self, *, states=None, observations=None) -> float:
        """Q = sum_i log pi_{a_i} + ll[a_i, i]  (MMM:579-624)."""
        return self._score(states, observations)[0]

    def model_log_likelihood(self, *, states=None, observations=None):
        """sum_i log sum_c pi_c exp(ll_ci)  (MMM:626-666)."""
        return self._score(states, observations)[1]

    def aic(self, states=None, observations=None) -> float:
        """MMM:668-693."""
        return -2 * self.model_log_likelihood(states=states, observations=observations) + 2 * self.n_free_params

    def bic(self, states=None, observations=None) -> float:
        """MMM:695-722."""
        n = self.n_data if states is None else np.atleast_3d(states).shape[1]
        return (-2 * self.model_log_likelihood(states=states, observations=observations)
                + np.log(n) * self.n_free_params)

    # ---- predictions built on the assignment probabilities (host arithmetic, O(n d)) ----------
    def one_step_ahead_predictions(self, *, states, observations):
        """MMM:808-850."""
        probs = self.mle_cluster_assignment(states=states, observations=observations, return_probs=True)[1]
        return self._mix_next(probs, np.atleast_3d(states), np.atleast_3d(observations))

    def one_step_ahead_predictions_no_history(self, *, states, observations):
        """MMM:852-907."""
        states, observations = np.atleast_3d(states), np.atleast_3d(observations)
        s0 = np.nan * np.ones_like(states)
        s0[-1] = states[-1]
        o0 = np.nan * np.ones_like(observations)
        o0[-1] = observations[-1]
        probs = self.mle_cluster_assignment(states=s0, observations=o0, return_probs=True)[1]
        return self._mix_next(probs, states, observations)

    def _mix_next(self, probs, states, observations):
        nxt_s = np.zeros(shape=(1, *states.shape[1:]))
        nxt_o = np.zeros(shape=(1, *observations.shape[1:]))
        last = states[-1]
        assert probs.shape == (self.n_clusters, last.shape[0])
        for c in range(self.n_clusters):
            s_c = last @ self.transition_matrices[c]
            o_c = s_c @ self.measurement_matrices[c]
            nxt_s += np.expand_dims(probs[c], axis=-1) * s_c
            nxt_o += np.expand_dims(probs[c], axis=-1) * o_c
        return nxt_s, nxt_o

    def predictions_from_initial_data(self, *, states=None, observations=None):
        """MMM:943-995 (predicts from the TRAINING set's first states, like the reference)."""
        a = self.initial_full_data_cluster_assignment(states=states, observations=observations)
        ps = np.zeros_like(self.states if states is None else states)
        po = np.zeros_like(self.observations if observations is None else observations)
        for i in range(self.n_data):
            z = np.atleast_1d(self.states[0, i, :])
            A = np.atleast_2d(self.transition_matrices[a[i]])
            H = np.atleast_2d(self.measurement_matrices[a[i]])
            for t in range(ps.shape[0]):
                ps[t, i, :] = z
                if t < po.shape[0]:
                    po[t, i, :] = z @ H
                z = z @ A
            for t in range(ps.shape[0], po.shape[0]):
                po[t, i, :] = z @ H
                z = z @ A
        return ps, po

    # ------------------------------------------------------------------ EM
    @staticmethod
    def regress(input_exogenous, output_endogenous):
        """Least-squares coefficients and residual covariance (MMM:1167-1200) from the moment
        kernel: one pseudo-cluster holding every row."""
        return MMLinGaussSS_marginalizable.regress_alpha(input_exogenous, output_endogenous, 0.0)

    @staticmethod
    def regress_alpha(input_exogenous, output_endogenous, alpha: float):
        """Ridge coefficients and residual covariance (MMM:1202-1233) on the device: the rows are
        presented to the M-step kernels as (z_t, x_t) pairs of a single cluster."""
        X = np.atleast_2d(np.asarray(input_exogenous, dtype=np.float64))
        Y = np.atleast_2d(np.asarray(output_endogenous, dtype=np.float64))
        if not (np.isfinite(X).all() and np.isfinite(Y).all()):
            # the moment kernels skip incomplete rows (M_step semantics); the reference's regress gets the rows as
            # they are and fails on them: numba's lstsq (MMM:1194) / sklearn's input validation (MMM:1226-1227)
            if alpha > 2 * np_eps:
                raise ValueError("Input X contains NaN.")
            raise np.linalg.LinAlgError("Array must not contain infs or NaNs.")
        if X.shape[1] > C.MTM_MAX_DIM or Y.shape[1] > C.MTM_MAX_DIM:
            raise C.MtmError(f"regress: at most {C.MTM_MAX_DIM} regressors / responses on the device path")
        eng = Engine.get_local()
        data = DeviceData(X[None], Y[None], eng.device)
        assign = torch.zeros(X.shape[0], dtype=torch.int32, device=eng.device)
        params, _, status = eng.mstep(data, assign, 1, alpha if alpha > 2 * np_eps else 0.0)
        _raise_for_status(status & ~C.STATUS_EMPTY_FIT if X.shape[0] else status, alpha)
        p = params.cpu().numpy()[0]
        d, l = X.shape[1], Y.shape[1]
        o = 1 + d + 3 * d * d
        return p[o:o + d * l].reshape(d, l).copy(), p[o + d * l:o + d * l + l * l].reshape(l, l).copy()

    def E_step(self) -> int:
        """New hard assignment; returns the number of switches (MMM:1235-1252)."""
        eng = self.engine
        assign = self._assign_dev()
        res = eng.estep(self._train_data(), self._packed_params(), self.n_clusters, self.n_timesteps, assign=assign,
                        count_switches=True)
        status, n_sw = res["status"], int(res["n_switches"])
        if self._world > 1:
            status = self._or_status(status)
            n_sw = mdist.sum_int(n_sw, eng.device, self._group if self._group is not False else None)
        _raise_for_status(status, self.alpha)
        self.cluster_assignment = self._assign_host(assign)
        return n_sw

    def M_step(self) -> None:
        """Parameters from the current hard assignment (MMM:1254-1308)."""
        params, _, status = self.engine.mstep(self._train_data(), self._assign_dev(), self.n_clusters, self.alpha)
        self._unpack_params(params)
        _raise_for_status(status, self.alpha)

    def train(self, *, verbose: bool = False, n_steps: int = 1000):
        """EM until no assignment changes, a cluster gets near-empty, or n_steps (MMM:1310-1366).
        The loop is resident on the device; only the control block comes back."""
        eng = self.engine
        assign = self._assign_dev()
        params, info = eng.em_fit(self._train_data(), assign, self.n_clusters, self.alpha, n_steps)
        self.fit_info = info
        self.cluster_assignment = self._assign_host(assign)
        if info["exit"] == "guard0":
            if verbose:
                print("Encountered near-empty cluster.")
            return self  # like the reference: returns before any M-step, last_trained stays None
        self._unpack_params(params)
        if info["exit"] == "error" or info["status"]:
            _raise_for_status(info["status"], self.alpha)
        if verbose:
            if info["exit"] == "converged":
                print(f"Optimisation completed in {info['n_e_steps'] - 1} steps.")
            elif info["exit"] == "guard":
                print("Encountered near-empty cluster.")
        self.last_trained = (datetime.datetime.now(datetime.timezone.utc).replace(microsecond=0)
                             .astimezone().isoformat())
        return self

    def train_with_multiple_random_starts(self, *, n_starts: int = 10, verbose: bool = False, n_steps: int = 100,
                                          return_objectives: bool = False, use_cache: bool = True):
        """k-means start (seed 0) + n_starts random starts (seeds 100+i); best by strict > on Q
        (MMM:1368-1471).  The training set is uploaded once and shared by all restarts."""
        if use_cache:
            cached = self._cached_model(verbose)
            if cached is not None:
                return cached

        shared = self._train_data()
        best_mdl = MMLinGaussSS_marginalizable(n_clusters=self.n_clusters, states=self.states,
                                               observations=self.observations, random_seed=0, init="kmeans",
                                               alpha=self.alpha, device=self._device, group=self._group,
                                               _shared_data=shared)
        try:
            best_mdl = best_mdl.train(verbose=verbose, n_steps=n_steps)
        except Exception:  # noqa: BLE001 - the reference swallows every failure of a start
            pass
        if return_objectives:
            objective_list = [best_mdl.e_complete_data_log_lik()]
        # The n_starts random restarts are independent given the data: they run as ONE batched fit
        # (mtm_sweep_fit: every kernel of an EM iteration serves all still-running restarts), then the
        # reference's sequential selection is replayed on the results.
        restarts = self._batched_random_restarts(shared, n_starts, n_steps) if n_starts > 0 else []
        for i in range(n_starts):
            try:
                status, q_i, make = restarts[i]
                _raise_for_status(status, self.alpha)  # a failed start raises inside train() in the reference
                if return_objectives:
                    objective_list.append(q_i)
                if q_i > best_mdl.e_complete_data_log_lik():
                    best_mdl = make()
            except Exception:  # noqa: BLE001,S110
                pass
        if not np.isfinite(best_mdl.e_complete_data_log_lik()):
            raise Exception("training failed")
        if use_cache and self._rank == 0:  # sharded: every rank holds the same model, one of them writes it
            best_mdl.to_pickle()
        if return_objectives:
            return best_mdl, np.array(objective_list)
        return best_mdl

    def _cached_model(self, verbose: bool = False):
        """The newest model cached for this data / n_clusters / alpha under ``<home>/tmp`` (file names
        ``mmm-<hex_hash>-<UTC stamp>.p.gz``, MMM:227-245), or None.  A file that cannot be loaded, or whose stored key
        differs, counts as a miss (MMM:1400-1428)."""
        hits = glob.glob(os.path.join(home_dir, "tmp", f"mmm-{self.hex_hash}*"))
        if not hits:
            if verbose:
                print("No model found in cache.")
            return None
        newest = max(hits, key=os.path.getmtime)
        try:
            mdl = MMLinGaussSS_marginalizable.from_pickle(
                newest, training_data={"states": self.states, "observations": self.observations})
            if mdl.hex_hash != self.hex_hash:
                raise ValueError("cache key mismatch")
        except Exception as err:  # noqa: BLE001
            if verbose:
                print(f"Issue loading cached model -- encountered {err}")
            return None
        if verbose:
            print(f"Loaded model {mdl.last_trained=} from cache.")
        return mdl

    def _batched_random_restarts(self, shared, n_starts, n_steps):
        """[(status, Q, factory)] for seeds 100 .. 100+n_starts-1 (MMM:1449-1455), fitted together."""
        eng = self.engine
        n, k = self.n_data, self.n_clusters
        lo, hi = (shared.shard[1], shared.shard[2]) if getattr(shared, "shard", None) else (0, n)
        if n >= 200_000:
            # large data: draw rng.integers(0, k, n) for every seed ON the device (bit-identical to numpy,
            # mtm_random_labels) instead of n_starts host draws + uploads; a sharded model keeps its rows of it
            assign = torch.empty((n_starts, hi - lo), dtype=torch.int32, device=eng.device)
            full = torch.empty(n, dtype=torch.int32, device=eng.device) if (lo, hi) != (0, n) else None
            for i in range(n_starts):
                if full is None:
                    eng.random_labels(100 + i, k, n, out=assign[i])
                else:
                    eng.random_labels(100 + i, k, n, out=full)
                    assign[i].copy_(full[lo:hi])
        else:
            a0 = np.stack([np.random.default_rng(seed=100 + i).integers(low=0, high=k, size=n)[lo:hi]
                           for i in range(n_starts)]).astype(np.int32)
            assign = torch.from_numpy(a0).to(eng.device)
        assign = shared.to_resident(assign)
        params, infos, scores, sstat = eng.sweep_fit(shared, assign, k, self.alpha, n_steps)
        params_h = params.cpu().numpy()
        assign_h = self._gather(shared, shared.from_resident(assign)).cpu().numpy()
        out = []
        for i in range(n_starts):
            info = infos[i]

            def make(i=i, info=info):
                mdl = MMLinGaussSS_marginalizable(n_clusters=k, states=self.states, observations=self.observations,
                                                  random_seed=100 + i, alpha=self.alpha, device=self._device,
                                                  group=self._group, _shared_data=shared)
                mdl.cluster_assignment = assign_h[i].astype(np.int64)
                mdl.fit_info = info
                if info["exit"] != "guard0":
                    mdl._unpack_params(params_h[i])
                    mdl.last_trained = (datetime.datetime.now(datetime.timezone.utc).replace(microsecond=0)
                                        .astimezone().isoformat())
                return mdl

            status = info["status"] | int(sstat[i])
            # a start that never trained (near-empty cluster at once) keeps the constructor's unseeded random
            # parameters in the reference; its Q is irreproducible noise there and NaN (never selected) here
            q_i = float(scores[i, 0]) if info["exit"] != "guard0" else float("nan")
            out.append((status, q_i, make))
        return out

    # ------------------------------------------------------------------ persistence (MMM:227-319)
    def to_pickle(self, save_location=None, there_can_only_be_one: bool = True, include_training_data: bool = False):
        save_location = os.path.join(home_dir, "tmp") if save_location is None else save_location
        os.makedirs(save_location, exist_ok=True)
        ts = datetime.datetime.now(datetime.timezone.utc).strftime("%Y%m%dT%H%MZ")
        if there_can_only_be_one:
            for f in glob.glob(os.path.join(save_location, f"mmm-{self.hex_hash}*")):
                os.remove(f)
        payload = {
            "n_clusters": self.n_clusters,
            "cluster_propensities": self.cluster_propensities,
            "init_state_means": self.init_state_means,
            "init_state_covs": self.init_state_covs,
            "transition_matrices": self.transition_matrices,
            "transition_covs": self.transition_covs,
            "measurement_matrices": self.measurement_matrices,
            "measurement_covs": self.measurement_covs,
            "random_seed": self.random_seed,
            "rng": self.rng,
            "init": self.init,
            "alpha": self.alpha,
            "cluster_assignment": self.cluster_assignment,
            "correspondence": self.correspondence,
            "inverse_correspondence": self.inverse_correspondence,
            "hex_hash": self.hex_hash,
            "time_stamp": self.time_stamp,
            "last_trained": self.last_trained,
        }
        if include_training_data:
            payload |= {"states": self.states, "observations": self.observations}
        with gzip.open(os.path.join(save_location, f"mmm-{self.hex_hash}-{ts}.p.gz"), "wb") as f:
            pickle.dump(payload, f)

    @staticmethod
    def from_pickle(file, training_data: dict = None):
        opener = gzip.open if os.path.splitext(file)[-1] == ".gz" else open
        with opener(file, "rb") as f:
            md = pickle.load(f)
        src = training_data if training_data is not None else md
        mdl = MMLinGaussSS_marginalizable(n_clusters=md["n_clusters"], states=src["states"],
                                          observations=src["observations"], random_seed=md["random_seed"],
                                          init=md["init"], alpha=md["alpha"] if "alpha" in md else 0)
        for key in ("cluster_propensities", "init_state_means", "init_state_covs", "transition_matrices",
                    "transition_covs", "measurement_matrices", "measurement_covs", "rng", "cluster_assignment"):
            setattr(mdl, key, md[key])
        mdl.correspondence = md["correspondence"]
        mdl.inverse_correspondence = md["inverse_correspondence"]
        mdl.time_stamp = md["time_stamp"]
        mdl.last_trained = md["last_trained"]
        return mdl
