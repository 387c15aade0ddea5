"""``StateSpaceModelClassifier``: generative (Bayes) classifier whose class-conditional densities are state-space
models — the interface of the reference's ``framework_extended/state_space_model_classifier.py`` (constructor,
``fit`` / ``score`` / ``predict_proba`` / ``predict``).

With linear-Gaussian components nothing loops over classes: ``fit`` is one ``mtm_mstep_ext`` call with the class
index as the label (all class models from one pass over the data), and every likelihood matrix is one E-step launch
(``_scoring.score_matrix``).  Other component types are fitted and scored one by one through their own methods."""

from __future__ import annotations

import numpy as np
import torch

from .. import _capi as C
from ..engine import DeviceData, Engine
from ._scoring import score_matrix
from .state_space_model_linear_gaussian import StateSpaceLinearGaussian, check_status


class StateSpaceModelClassifier:
    def __init__(self, component_model):
        self.component_model = component_model
        self.classes = self.n_classes = self.propensities = self.class_models = self.data = None

    def fit(self, data, labels):
        self.data = tuple(map(np.atleast_3d, data))
        states, measurements = self.data
        labels = np.asarray(labels)
        self.classes, index, counts = np.unique(labels, return_inverse=True, return_counts=True)
        self.n_classes = len(self.classes)
        self.propensities = counts / counts.sum()
        self.class_models = [self.component_model() for _ in self.classes]
        members = [np.flatnonzero(index == k) for k in range(self.n_classes)]
        if self.component_model is StateSpaceLinearGaussian:
            eng = Engine.get_local()
            dd = DeviceData(states, measurements, eng.device)
            lab = torch.from_numpy(index.astype(np.int32)).to(eng.device)
            params, _, status = eng.mstep_ext(dd, lab, self.n_classes, self.class_models[0].alpha)
            check_status(status & ~C.STATUS_EMPTY_FIT)
            rows = params.cpu().numpy()
            for k, mdl in enumerate(self.class_models):
                mdl._adopt(rows[k], dd.d, dd.l).data = (states[:, members[k]], measurements[:, members[k]])
        else:
            for k, mdl in enumerate(self.class_models):
                mdl.fit(data=(states[:, members[k]], measurements[:, members[k]]))
        return self

    def _joint(self, data):
        """(n, n_classes) matrix of p_c * p(data_i | c) and the data it was evaluated on"""
        data = self.data if data is None else tuple(map(np.atleast_3d, data))
        with np.errstate(all="ignore"):
            joint = np.exp(score_matrix(self.class_models, data)).T * self.propensities[None, :]
        assert joint.shape[0] == data[0].shape[1]
        return joint

    def score(self, data=None) -> float:
        """log-likelihood of the data under the class mixture: sum_i log sum_c p_c p(data_i | c)"""
        with np.errstate(all="ignore"):
            return float(np.log(self._joint(data).sum(axis=1)).sum())

    def predict_proba(self, data=None) -> np.ndarray:
        joint = self._joint(data)
        post = joint / joint.sum(axis=1, keepdims=True)
        assert np.all(post >= 0.0) and np.allclose(post.sum(axis=-1), 1.0)
        return post

    def predict(self, data=None) -> np.ndarray:
        return self.classes[self.predict_proba(data).argmax(axis=1)]
