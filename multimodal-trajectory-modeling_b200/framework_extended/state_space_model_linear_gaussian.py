"""Drop-in for the reference's ``framework_extended/state_space_model_linear_gaussian.py``:
``StateSpaceLinearGaussian`` — ONE linear-Gaussian state-space model, the pluggable component of
``StateSpaceMixtureModel`` / ``StateSpaceModelClassifier``.

* ``score`` (reference :99-126) is the hot function — the same stacked-mean / dense-covariance /
  marginalising log-density call as the mixture's E-step — and runs on the E-step kernel with a single
  cluster (pi = 1).
* ``fit`` (reference :39-79) is a one-shot host computation with its own semantics, different from the
  mixture's M-step: initial mean = ``np.nanmean`` per coordinate, initial covariance from the fully finite
  rows, regressions by ``util.regress`` (normal equations with a fixed 1e-6 ridge, solved by lstsq;
  util_state_space.py:33-61) or ``util.regress_alpha`` (sklearn Ridge, :64-94).  Restated here in numpy.
"""

from __future__ import annotations

import pickle

import numpy as np

from ..engine import DeviceData, Engine

np_eps = np.finfo(float).eps


def _regress(X, Y, alpha):
    """coefficients and residual covariance on the rows where X and Y are completely finite"""
    ok = np.isfinite(np.column_stack((X, Y))).all(axis=-1)
    X, Y = X[ok], Y[ok]
    p = X.shape[1]
    if alpha > 2 * np_eps:  # Ridge(alpha, fit_intercept=False): (X'X + alpha I) B = X'Y
        B = np.linalg.solve(X.T @ X + alpha * np.eye(p), X.T @ Y)
    else:  # util.regress: lstsq on the normal equations with eps = 1e-6
        B = np.linalg.lstsq(X.T @ X + 1e-6 * np.eye(p), X.T @ Y, rcond=-1)[0]
    return B, np.cov(Y - X @ B, rowvar=False)


class StateSpaceLinearGaussian:
    """Linear-Gaussian state-space model with the reference's estimator-style interface."""

    def __init__(self, alpha: float = 0.0, *, device=None):
        self.alpha = alpha if alpha > 2 * np_eps else 0
        self.data = None
        self._device = device

    def __str__(self):
        return "State space model with linear Gaussian components"

    def fit(self, data):
        self.data = tuple(map(np.atleast_3d, data))
        states, measurements = self.data
        d, l = states.shape[2], measurements.shape[2]
        z0 = states[0]
        with np.errstate(all="ignore"):
            self.state_init = {"mean": np.nanmean(z0, axis=0),
                               "cov": np.cov(z0[np.isfinite(z0).all(axis=1)], rowvar=False)}
            self.state_model = dict(zip(["coeff", "covar"],
                                        _regress(states[:-1].reshape(-1, d), states[1:].reshape(-1, d), self.alpha)))
            self.measurement_model = dict(zip(["coeff", "covar"],
                                              _regress(states.reshape(-1, d), measurements.reshape(-1, l), self.alpha)))
        return self

    def to_pickle(self) -> bytes:
        return pickle.dumps({"state_init": self.state_init, "state_model": self.state_model,
                             "measurement_model": self.measurement_model, "alpha": self.alpha})

    def from_pickle(self, p: bytes):
        d = pickle.loads(p)
        self.state_init, self.state_model, self.measurement_model = d["state_init"], d["state_model"], d["measurement_model"]
        self.alpha = d["alpha"] if "alpha" in d else 0
        return self

    def _packed(self, d, l):
        return np.concatenate([[1.0], np.reshape(self.state_init["mean"], d), np.reshape(self.state_init["cov"], d * d),
                               np.reshape(self.state_model["coeff"], d * d), np.reshape(self.state_model["covar"], d * d),
                               np.reshape(self.measurement_model["coeff"], d * l),
                               np.reshape(self.measurement_model["covar"], l * l)])[None]

    def score(self, data=None):
        """log p(z_i, x_i) per trajectory, missing entries marginalised (reference :99-126)."""
        states, measurements = map(np.atleast_3d, self.data if data is None else data)
        eng = Engine.get(self._device)
        dd = DeviceData(states, measurements, eng.device)
        res = eng.estep(dd, self._packed(dd.d, dd.l), 1, dd.T, want_ll=True)
        if res["status"] & 1:
            raise np.linalg.LinAlgError("Array must not contain infs or NaNs.")
        return res["ll"][0].cpu().numpy()
