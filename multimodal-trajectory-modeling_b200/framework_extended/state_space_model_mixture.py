"""Drop-in for the reference's ``framework_extended/state_space_model_mixture.py``: ``StateSpaceMixtureModel``,
the hard-assignment EM mixture with pluggable component models.  Same constructor / ``fit`` / ``predict_proba`` /
``predict`` / ``score`` / ``model_log_likelihood`` / ``cluster_assignment_index`` / pickling and the same control
flow (guards, swallowed failures, restart selection).  With linear-Gaussian components the E-step likelihoods of
all clusters come from ONE launch of the E-step kernel on a training set that is uploaded once
(``_scoring.score_matrix``); the M-step keeps the component's own ``fit`` semantics (nanmean / 1e-6-ridge
regressions of ``util.regress``), which differ from ``MMLinGaussSS_marginalizable.M_step``."""

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
from ._scoring import score_matrix
from .state_space_model_linear_gaussian import StateSpaceLinearGaussian

home_dir = os.environ.get("MTM_B200_HOME", os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def take_finite_along_axis(arr, axis: int = 0):
    """slices along `axis` whose every element is finite (util_state_space.py:342-383)"""
    new_shape = list(arr.shape)
    new_shape[axis] = -1
    keep = np.argwhere(np.isfinite(arr).all(axis=tuple(a for a in range(arr.ndim) if a != axis)))
    return np.take(arr, keep, axis).reshape(tuple(new_shape))


class StateSpaceMixtureModel:
    """Mixture of state space models"""

    # the default generator is created ONCE, at definition time, and shared by every instance that does not
    # pass its own — exactly like the reference's signature (state_space_model_mixture.py:39)
    def __init__(self, n_clusters: int, data, component_model, *, component_model_hyperparams: dict = dict(),
                 rng: np.random.Generator = np.random.default_rng(seed=42)):
        self.rng = rng
        self.states, self.observations = map(np.atleast_3d, data)
        self.n_timesteps, self.n_data, self.d_states = self.states.shape
        self.d_observations = self.observations.shape[-1]
        self.n_clusters = n_clusters
        self.cluster_propensities = np.ones(self.n_clusters) / self.n_clusters
        self.cluster_assignment = self.rng.integers(self.n_clusters, size=self.n_data)
        self.component_model = component_model
        self.component_model_hyperparams = component_model_hyperparams
        self.cluster_models = [self.component_model(**self.component_model_hyperparams) for _ in range(self.n_clusters)]
        self._correspondence = dict(zip(range(self.n_clusters), string.ascii_uppercase))
        self.inverse_correspondence = {v: k for k, v in self._correspondence.items()}
        self.hex_hash = hashlib.md5(
            self.states.tobytes() + self.observations.tobytes() + str(self.n_clusters).encode("utf-8")
            + str(self.component_model).encode("utf-8")
            + (json.dumps(self.component_model_hyperparams, sort_keys=True).encode("utf-8")
               if self.component_model_hyperparams != {} else b"")).hexdigest()
        self.time_stamp = datetime.datetime.now(datetime.timezone.utc).replace(microsecond=0).astimezone().isoformat()
        self.last_trained = None
        self._resident = None  # training set on the device (linear-Gaussian components only)

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
        self.inverse_correspondence = {v: k for k, v in self._correspondence.items()}

    # ------------------------------------------------------------------ likelihoods
    def _ll(self, data=None) -> np.ndarray:
        """(n_clusters, n) conditional log-likelihoods; the training set stays resident on the device"""
        if data is None:
            if self.component_model is StateSpaceLinearGaussian:
                if self._resident is None:
                    self._resident = DeviceData(self.states, self.observations, Engine.get().device)
                return score_matrix(self.cluster_models, self._resident)
            data = self.data
        return score_matrix(self.cluster_models, data)

    def _E_step(self):
        """hard assignment of every training point to its most likely cluster (reference :110-132)"""
        with np.errstate(all="ignore"):
            new_cluster_assignment = np.argmax((self.cluster_propensities[:, None] * np.exp(self._ll())).T, axis=1)
        assert new_cluster_assignment.size == self.n_data
        assert set(new_cluster_assignment) == set(range(self.n_clusters))
        n_switches = int(np.sum(np.not_equal(self.cluster_assignment, new_cluster_assignment)))
        self.cluster_assignment = new_cluster_assignment
        return n_switches

    def _M_step(self):
        """component-wise refit on the members of every cluster (reference :134-149)"""
        for c in range(self.n_clusters):
            self.cluster_propensities[c] = np.mean(self.cluster_assignment == c)
            self.cluster_models[c].fit((self.states[:, self.cluster_assignment == c],
                                        self.observations[:, self.cluster_assignment == c]))
        assert np.isclose(sum(self.cluster_propensities), 1.0)

    def fit(self, *, init: str = "random", n_iter: int = 1000, n_restarts: int = 0, use_cache: bool = True,
            verbose: bool = False):
        """EM with the reference's initialisations, guards and restart selection (reference :151-308)"""
        from sklearn import cluster as skl_cluster

        if bool(use_cache):
            try:
                pfile = sorted(glob.glob(os.path.join(home_dir, "tmp", f"mmm-{self.hex_hash}*")),
                               key=os.path.getmtime).pop()
                best_mdl = StateSpaceMixtureModel.from_pickle(
                    pfile, training_data={"states": self.states, "observations": self.observations})
                assert self.hex_hash == best_mdl.hex_hash
                if verbose:
                    print(f"Loaded model {best_mdl.last_trained=} from cache.")
                return best_mdl
            except IndexError:
                if verbose:
                    print("No model found in cache.")
            except AssertionError:
                if verbose:
                    print("Model found in cache does not match our requirements.")
            except Exception as err:  # noqa: BLE001 - mirrors the reference
                if verbose:
                    print(f"Issue loading cached model -- encountered {err}")

        km = lambda X: skl_cluster.KMeans(n_clusters=self.n_clusters, init="k-means++", random_state=0).fit_predict(X)
        if init in ("k-means", "kmeans"):
            self.cluster_assignment = km(self.states[0])
        elif init in ("kmeans-all", "k-means-all"):
            self.cluster_assignment = km(np.vstack([self.states[:, i, :].flatten() for i in range(self.n_data)]))
        elif init in ("kmeans-take-finite", "k-means-take-finite"):
            self.cluster_assignment = km(np.column_stack(take_finite_along_axis(self.states, 0)))
        else:
            self.cluster_assignment = self.rng.integers(low=0, high=self.n_clusters, size=self.n_data)
        assert len(self.cluster_assignment) == self.n_data
        if np.min(np.bincount(self.cluster_assignment)) <= 3:
            warnings.warn("Cluster initialisation method yielded a nearlyempty cluster")
            self.cluster_assignment = self.rng.integers(low=0, high=self.n_clusters, size=self.n_data)

        try:
            self._M_step()
            for _ in range(n_iter):
                n_switches = self._E_step()
                if n_switches == 0:
                    break
                if np.min(np.bincount(self.cluster_assignment)) <= 3:
                    raise Exception("Encountered nearly empty cluster.")
                self._M_step()
        except Exception:  # noqa: BLE001 - the reference swallows every failure of a fit
            pass

        try:
            score = self.score()
        except TypeError:
            score = -np.inf
        best_mdl, best_score = self, score
        for i in range(n_restarts):
            try:
                new_candidate = StateSpaceMixtureModel(
                    n_clusters=self.n_clusters, data=self.data, component_model=self.component_model,
                    component_model_hyperparams=self.component_model_hyperparams,
                    rng=np.random.default_rng(seed=i)).fit(init="random", n_iter=n_iter)
                if (new_score := new_candidate.score()) > best_score:
                    best_mdl, best_score = new_candidate, new_score
            except Exception:  # noqa: BLE001,S110
                pass
        if best_score == -np.inf:
            raise Exception("training failed")
        best_mdl.last_trained = (datetime.datetime.now(datetime.timezone.utc).replace(microsecond=0)
                                 .astimezone().isoformat())
        if use_cache:
            best_mdl.to_pickle(include_training_data=False)
        return best_mdl

    # ------------------------------------------------------------------ persistence (reference :310-404)
    def to_pickle(self, save_location=None, there_can_only_be_one: bool = True, include_training_data: bool = False):
        save_location = os.path.join(home_dir, "tmp") if save_location is None else save_location
        os.makedirs(save_location, exist_ok=True)
        ts = datetime.datetime.now(datetime.timezone.utc).strftime("%Y%m%dT%H%MZ")
        if there_can_only_be_one:
            for f in glob.glob(os.path.join(save_location, f"mmm-{self.hex_hash}*")):
                os.remove(f)
        with gzip.open(os.path.join(save_location, f"mmm-{self.hex_hash}-{ts}.p.gz"), "wb") as f:
            pickle.dump({"n_clusters": self.n_clusters, "cluster_propensities": self.cluster_propensities,
                         "component_model": self.component_model,
                         "component_model_hyperparams": self.component_model_hyperparams,
                         "cluster_models": [cm.to_pickle() for cm in self.cluster_models], "rng": self.rng,
                         "cluster_assignment": self.cluster_assignment, "correspondence": self.correspondence,
                         "inverse_correspondence": self.inverse_correspondence, "hex_hash": self.hex_hash,
                         "time_stamp": self.time_stamp, "last_trained": self.last_trained}
                        | ({"states": self.states, "observations": self.observations} if include_training_data
                           else {}), f)

    @staticmethod
    def from_pickle(file, training_data: dict = None):
        with (gzip.open(file, "rb") if os.path.splitext(file)[-1] == ".gz" else open(file, "rb")) as f:
            d = pickle.load(f)
        src = training_data if training_data is not None else d
        mdl = StateSpaceMixtureModel(n_clusters=d["n_clusters"], data=(src["states"], src["observations"]),
                                     component_model=d["component_model"],
                                     component_model_hyperparams=d.get("component_model_hyperparams", dict()),
                                     rng=d["rng"])
        mdl.cluster_propensities = d["cluster_propensities"]
        mdl.cluster_models = [mdl.component_model().from_pickle(p) for p in d["cluster_models"]]
        mdl.rng = d["rng"]
        mdl.cluster_assignment = d["cluster_assignment"]
        mdl.correspondence = d["correspondence"]
        mdl.inverse_correspondence = d["inverse_correspondence"]
        mdl.last_trained = d["last_trained"]
        return mdl

    # ------------------------------------------------------------------ predictions (reference :406-512)
    def predict_proba(self, data=None, return_prenormalized_log_probs: bool = False):
        ll = self._ll(data)
        with np.errstate(all="ignore"):
            preds = (self.cluster_propensities[:, None] * np.exp(ll)).T
            preds /= preds.sum(axis=1, keepdims=True)
            if return_prenormalized_log_probs:
                return preds, np.log(self.cluster_propensities)[:, None] + ll
        return preds

    def predict(self, *, data=None, letters: bool = True):
        preds = np.argmax(self.predict_proba(data=data), axis=1)
        return np.array([self.correspondence[i] for i in preds]) if letters else preds

    def score(self, data=None) -> float:
        """expected complete-data log-likelihood under the hard assignment; -inf if a cluster is empty"""
        n = self.n_data if data is None else np.atleast_3d(data[0]).shape[1]
        ll = self._ll(data)
        with np.errstate(all="ignore"):
            preds = (self.cluster_propensities[:, None] * np.exp(ll)).T
            a = np.argmax(preds / preds.sum(axis=1, keepdims=True), axis=1)  # predict(): argmax of the normalised rows
        if set(a) != set(range(self.n_clusters)) or a.size != n:
            return -np.inf
        with np.errstate(all="ignore"):
            return float(np.sum(np.log(self.cluster_propensities[a])) + np.sum(ll[a, np.arange(a.size)]))

    def model_log_likelihood(self, data=None) -> float:
        with np.errstate(all="ignore"):
            return float(np.sum(np.log(np.sum(self.cluster_propensities[:, None] * np.exp(self._ll(data)), axis=0))))

    def cluster_assignment_index(self, *, cluster: str = "A", data=None):
        return self.predict_proba(data=data, return_prenormalized_log_probs=True)[-1][self.inverse_correspondence[cluster]]
