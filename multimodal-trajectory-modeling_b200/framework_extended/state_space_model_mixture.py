"""``StateSpaceMixtureModel``: hard-assignment EM mixture of pluggable state-space component models — the interface
of the reference's ``framework_extended/state_space_model_mixture.py`` (SSMM): constructor, ``fit``, ``predict_proba``,
``predict``, ``score``, ``model_log_likelihood``, ``cluster_assignment_index``, ``to_pickle`` / ``from_pickle`` with the
same payload keys and cache file names.

With linear-Gaussian components the whole fit is resident on the device (``mtm_em_fit_ext``): the training set is
uploaded once, every iteration is the fused E-step + statistics kernel plus a parameter solve with the component's
own estimation rules (nanmean initial mean, 1e-6-ridge regressions), and the loop applies SSMM.fit's stopping rules
on the device; only the final labels and parameter rows come back.  Every likelihood matrix of ``predict*`` /
``score`` is one E-step launch.  Any other component class is driven through its own ``fit`` / ``score`` on the
host, one cluster at a time, with the same stopping rules.

Behaviour kept from the reference because callers can observe it: the default generator is created once, at
definition time; the constructor draws an assignment from it; a near-empty initial assignment is re-drawn, a
near-empty cluster later ends the fit silently; restart candidates are fitted through the cache (each candidate's
``fit`` runs with ``use_cache=True``), the best finite score wins."""

from __future__ import annotations

import datetime
import glob
import gzip
import hashlib
import json
import os
import pickle
import string
import warnings

import numpy as np

from ..engine import DeviceData, Engine
from ._scoring import all_linear_gaussian, score_matrix
from .state_space_model_linear_gaussian import StateSpaceLinearGaussian

home_dir = os.environ.get("MTM_B200_HOME", os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

_KMEANS_INITS = {"k-means": "first", "kmeans": "first", "k-means-all": "all", "kmeans-all": "all",
                 "k-means-take-finite": "finite", "kmeans-take-finite": "finite"}


def _now():
    return datetime.datetime.now(datetime.timezone.utc).replace(microsecond=0).astimezone().isoformat()


def take_finite_along_axis(arr, axis: int = 0):
    """slices along `axis` whose every element is finite (util_state_space.py:342-383)"""
    others = tuple(a for a in range(arr.ndim) if a != axis)
    return np.compress(np.isfinite(arr).all(axis=others), arr, axis=axis)


class StateSpaceMixtureModel:
    """Mixture of state space models"""

    def __init__(self, n_clusters: int, data, component_model, *, component_model_hyperparams: dict = dict(),
                 rng: np.random.Generator = np.random.default_rng(seed=42)):
        self.rng = rng
        self.states, self.observations = map(np.atleast_3d, data)
        self.n_timesteps, self.n_data, self.d_states = self.states.shape
        self.d_observations = self.observations.shape[-1]
        self.n_clusters = n_clusters
        self.component_model = component_model
        self.component_model_hyperparams = component_model_hyperparams
        self.cluster_propensities = np.full(n_clusters, 1.0 / n_clusters)
        self.cluster_assignment = self.rng.integers(n_clusters, size=self.n_data)  # advances the generator (SSMM:52-54)
        self.cluster_models = [component_model(**component_model_hyperparams) for _ in range(n_clusters)]
        self.correspondence = dict(zip(range(n_clusters), string.ascii_uppercase))
        # cache key of SSMM:71-84: data bytes, n_clusters, the component class and its hyper-parameters
        key = hashlib.md5(self.states.tobytes() + self.observations.tobytes())
        key.update(str(n_clusters).encode("utf-8") + str(component_model).encode("utf-8"))
        if component_model_hyperparams != {}:
            key.update(json.dumps(component_model_hyperparams, sort_keys=True).encode("utf-8"))
        self.hex_hash = key.hexdigest()
        self.time_stamp = _now()
        self.last_trained = None
        self._resident = None  # the training set on the device

    def __str__(self):
        return "Mixture of state space models with {} components".format(self.component_model)

    @property
    def data(self):
        return self.states, self.observations

    @property
    def correspondence(self):
        return self._correspondence

    @correspondence.setter
    def correspondence(self, corr) -> None:
        self._correspondence = corr
        self.inverse_correspondence = {letter: index for index, letter in corr.items()}

    # ------------------------------------------------------------------ device plumbing
    @property
    def _on_device(self) -> bool:
        return all_linear_gaussian(self.cluster_models)

    def _train_data(self) -> DeviceData:
        if self._resident is None:
            self._resident = DeviceData(self.states, self.observations, Engine.get_local().device, pattern_order=True)
        return self._resident

    def _ll(self, data=None) -> np.ndarray:
        """(n_clusters, n) conditional log-likelihoods in the caller's order"""
        if data is None:
            data = self._train_data() if self._on_device else self.data
        return score_matrix(self.cluster_models, data)

    def _joint(self, data=None):
        """(n, n_clusters) pi_c p(data_i | c) and the log-likelihood matrix it came from"""
        ll = self._ll(data)
        with np.errstate(all="ignore"):
            return np.exp(ll).T * self.cluster_propensities[None, :], ll

    def _adopt_rows(self, rows):
        rows = np.asarray(rows)
        self.cluster_propensities[:] = rows[:, 0]
        for c, mdl in enumerate(self.cluster_models):
            mdl._adopt(rows[c], self.d_states, self.d_observations)
            mdl.data = (self.states[:, self.cluster_assignment == c], self.observations[:, self.cluster_assignment == c])

    # ------------------------------------------------------------------ EM
    def _M_step(self):
        """parameters of every cluster from its current members (SSMM:134-149)"""
        if self._on_device:
            data = self._train_data()
            lab = data.to_resident(np.ascontiguousarray(self.cluster_assignment, dtype=np.int32))
            rows, _, _ = Engine.get_local().mstep_ext(data, lab, self.n_clusters, self.cluster_models[0].alpha)
            self._adopt_rows(rows.cpu().numpy())
        else:
            for c, mdl in enumerate(self.cluster_models):
                member = self.cluster_assignment == c
                self.cluster_propensities[c] = member.mean()
                mdl.fit((self.states[:, member], self.observations[:, member]))
        assert np.isclose(self.cluster_propensities.sum(), 1.0)

    def _E_step(self) -> int:
        """most likely cluster of every training trajectory; raises before storing anything if a cluster would be
        left without members (SSMM:110-132)"""
        labels = self._joint()[0].argmax(axis=1)
        assert labels.size == self.n_data and np.unique(labels).size == self.n_clusters
        changed = int(np.count_nonzero(labels != self.cluster_assignment))
        self.cluster_assignment = labels
        return changed

    def _em_on_device(self, n_iter: int):
        eng, data = Engine.get_local(), self._train_data()
        lab = data.to_resident(np.ascontiguousarray(self.cluster_assignment, dtype=np.int32))
        rows, self.fit_info = eng.em_fit_ext(data, lab, self.n_clusters, self.cluster_models[0].alpha, n_iter)
        self.cluster_assignment = data.from_resident(lab).cpu().numpy().astype(np.int64)
        if self.fit_info["n_m_steps"] > 0:
            self._adopt_rows(rows.cpu().numpy())

    def _em_on_host(self, n_iter: int):
        try:  # any failure — a component's, an emptied cluster — ends the fit where it stands (SSMM:262-276)
            self._M_step()
            for _ in range(n_iter):
                if self._E_step() == 0:
                    return
                if np.bincount(self.cluster_assignment).min() <= 3:
                    return
                self._M_step()
        except Exception:  # noqa: BLE001
            return

    def _initial_labels(self, init: str):
        kind = _KMEANS_INITS.get(init)
        if kind is None:
            return self.rng.integers(low=0, high=self.n_clusters, size=self.n_data)
        from sklearn.cluster import KMeans

        if kind == "first":
            features = self.states[0]
        elif kind == "all":
            features = self.states.transpose(1, 0, 2).reshape(self.n_data, -1)
        else:
            features = np.column_stack(take_finite_along_axis(self.states, 0))
        return KMeans(n_clusters=self.n_clusters, init="k-means++", random_state=0).fit_predict(features)

    def _from_cache(self, verbose: bool):
        """newest cached model of this data / configuration, or None (SSMM:186-217)"""
        hits = sorted(glob.glob(os.path.join(home_dir, "tmp", f"mmm-{self.hex_hash}*")), key=os.path.getmtime)
        if not hits:
            if verbose:
                print("No model found in cache.")
            return None
        try:
            cached = StateSpaceMixtureModel.from_pickle(hits[-1], training_data={"states": self.states,
                                                                               "observations": self.observations})
        except Exception as err:  # noqa: BLE001
            if verbose:
                print(f"Issue loading cached model -- encountered {err}")
            return None
        if cached.hex_hash != self.hex_hash:
            if verbose:
                print("Model found in cache does not match our requirements.")
            return None
        if verbose:
            print(f"Loaded model {cached.last_trained=} from cache.")
        return cached

    def fit(self, *, init: str = "random", n_iter: int = 1000, n_restarts: int = 0, use_cache: bool = True,
            verbose: bool = False):
        """EM from the chosen initialisation, then `n_restarts` random restarts; returns the best model (SSMM:151-308)"""
        if use_cache:
            cached = self._from_cache(verbose)
            if cached is not None:
                return cached
        labels = self._initial_labels(init)
        assert len(labels) == self.n_data
        if np.bincount(labels).min() <= 3:
            warnings.warn("Cluster initialisation method yielded a nearlyempty cluster")
            labels = self.rng.integers(low=0, high=self.n_clusters, size=self.n_data)
        self.cluster_assignment = labels
        (self._em_on_device if self._on_device else self._em_on_host)(n_iter)

        def safe_score(mdl):
            try:
                return mdl.score()
            except TypeError:  # components that never got fitted
                return -np.inf

        best, best_score = self, safe_score(self)
        for seed in range(n_restarts):
            try:  # a candidate goes through the cache like any top-level fit (SSMM:285-291)
                cand = StateSpaceMixtureModel(self.n_clusters, self.data, self.component_model,
                                              component_model_hyperparams=self.component_model_hyperparams,
                                              rng=np.random.default_rng(seed=seed)).fit(init="random", n_iter=n_iter)
                cand_score = cand.score()
            except Exception:  # noqa: BLE001
                continue
            if cand_score > best_score:
                best, best_score = cand, cand_score
        if best_score == -np.inf:
            raise Exception("training failed")
        best.last_trained = _now()
        if use_cache:
            best.to_pickle(include_training_data=False)
        return best

    # ------------------------------------------------------------------ predictions
    def predict_proba(self, data=None, return_prenormalized_log_probs: bool = False):
        """(n, n_clusters) membership probabilities [, (n_clusters, n) log pi_c + log p(data_i | c)]"""
        joint, ll = self._joint(data)
        with np.errstate(all="ignore"):
            post = joint / joint.sum(axis=1, keepdims=True)
            if return_prenormalized_log_probs:
                return post, ll + np.log(self.cluster_propensities)[:, None]
        return post

    def predict(self, *, data=None, letters: bool = True):
        labels = self.predict_proba(data=data).argmax(axis=1)
        return np.array([self.correspondence[k] for k in labels]) if letters else labels

    def score(self, data=None) -> float:
        """expected complete-data log-likelihood under the hard assignment; -inf when a cluster attracts no point"""
        joint, ll = self._joint(data)
        with np.errstate(all="ignore"):
            labels = (joint / joint.sum(axis=1, keepdims=True)).argmax(axis=1)
            if not np.array_equal(np.unique(labels), np.arange(self.n_clusters)):
                return -np.inf
            return float(np.log(self.cluster_propensities[labels]).sum() + ll[labels, np.arange(labels.size)].sum())

    def model_log_likelihood(self, data=None) -> float:
        with np.errstate(all="ignore"):
            return float(np.log(self._joint(data)[0].sum(axis=1)).sum())

    def cluster_assignment_index(self, *, cluster: str = "A", data=None):
        """pre-normalised log-odds of assignment to `cluster`"""
        return self.predict_proba(data=data, return_prenormalized_log_probs=True)[1][self.inverse_correspondence[cluster]]

    # ------------------------------------------------------------------ persistence (payload keys of SSMM:337-357)
    def to_pickle(self, save_location=None, there_can_only_be_one: bool = True, include_training_data: bool = False):
        folder = os.path.join(home_dir, "tmp") if save_location is None else save_location
        os.makedirs(folder, exist_ok=True)
        if there_can_only_be_one:
            for old in glob.glob(os.path.join(folder, f"mmm-{self.hex_hash}*")):
                os.remove(old)
        payload = {key: getattr(self, key) for key in (
            "n_clusters", "cluster_propensities", "component_model", "component_model_hyperparams", "rng",
            "cluster_assignment", "correspondence", "inverse_correspondence", "hex_hash", "time_stamp", "last_trained")}
        payload["cluster_models"] = [mdl.to_pickle() for mdl in self.cluster_models]
        if include_training_data:
            payload.update(states=self.states, observations=self.observations)
        stamp = datetime.datetime.now(datetime.timezone.utc).strftime("%Y%m%dT%H%MZ")
        with gzip.open(os.path.join(folder, f"mmm-{self.hex_hash}-{stamp}.p.gz"), "wb") as f:
            pickle.dump(payload, f)

    @staticmethod
    def from_pickle(file, training_data: dict = None):
        opener = gzip.open if os.path.splitext(file)[-1] == ".gz" else open
        with opener(file, "rb") as f:
            payload = pickle.load(f)
        arrays = payload if training_data is None else training_data
        mdl = StateSpaceMixtureModel(payload["n_clusters"], (arrays["states"], arrays["observations"]),
                                     payload["component_model"],
                                     component_model_hyperparams=payload.get("component_model_hyperparams", dict()),
                                     rng=payload["rng"])
        mdl.cluster_models = [mdl.component_model().from_pickle(blob) for blob in payload["cluster_models"]]
        for key in ("cluster_propensities", "rng", "cluster_assignment", "correspondence", "last_trained"):
            setattr(mdl, key, payload[key])
        return mdl
