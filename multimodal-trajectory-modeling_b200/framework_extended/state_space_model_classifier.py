"""Drop-in for the reference's ``framework_extended/state_space_model_classifier.py``: a generative (Bayes)
classifier whose class-conditional densities are state-space models.  Same constructor, ``fit`` / ``score`` /
``predict_proba`` / ``predict``; with linear-Gaussian components all class likelihoods of a data set come from one
E-step launch (``_scoring.score_matrix``)."""

from __future__ import annotations

import numpy as np

from ._scoring import score_matrix


class StateSpaceModelClassifier:
    def __init__(self, component_model):
        self.component_model = component_model
        self.classes, self.n_classes = None, None
        self.propensities = None
        self.class_models = None
        self.data = None

    def fit(self, data, labels):
        """one component per class, fitted on that class's trajectories (reference :33-44)"""
        self.data = tuple(map(np.atleast_3d, data))
        states, measurements = data
        self.classes, cts = np.unique(labels, return_counts=True)
        self.n_classes = len(self.classes)
        self.propensities = cts / np.sum(cts)
        self.class_models = [self.component_model() for _ in self.classes]
        for i, c in enumerate(self.classes):
            self.class_models[i].fit(data=(states[:, labels == c], measurements[:, labels == c]))
        return self

    def _resolve(self, data):
        return self.data if data is None else tuple(map(np.atleast_3d, data))

    def _joint(self, data):
        with np.errstate(all="ignore"):
            return (self.propensities[:, None] * np.exp(score_matrix(self.class_models, data))).T  # (n, n_classes)

    def score(self, data=None) -> float:
        """sum_i log sum_c p_c p(data_i | c)  (reference :46-63)"""
        data = self._resolve(data)
        jt_p = np.sum(self._joint(data), axis=1)
        assert jt_p.shape[0] == data[0].shape[1]
        return float(np.sum(np.log(jt_p)))

    def predict_proba(self, data=None) -> np.ndarray:
        """(n, n_classes) posterior class probabilities (reference :65-84)"""
        data = self._resolve(data)
        pc = self._joint(data)
        pc /= np.sum(pc, axis=1, keepdims=True)
        assert pc.shape[0] == data[0].shape[1]
        assert np.all(pc >= 0.0) and np.allclose(np.sum(pc, axis=-1), 1.0)
        return pc

    def predict(self, data=None) -> np.ndarray:
        """reference :86-97"""
        data = self._resolve(data)
        preds = self.classes[np.argmax(self.predict_proba(data), axis=1)]
        assert preds.size == data[0].shape[1]
        return preds
