"""``StateSpaceLinearGaussian``: ONE linear-Gaussian state-space model with the estimator interface of the
reference's ``framework_extended/state_space_model_linear_gaussian.py`` (SSLG) — the pluggable component of
``StateSpaceMixtureModel`` / ``StateSpaceModelClassifier``.

Both halves run on the device.  ``score`` (SSLG:99-126: stacked mean, dense covariance, marginalising log-density)
is the E-step kernel with a single parameter row (pi = 1).  ``fit`` (SSLG:39-79) is ``mtm_mstep_ext`` with a single
label: the sufficient statistics of the data set are reduced on the device and solved with the component's own
rules — initial mean = np.nanmean per coordinate, covariance from the completely finite rows, regressions by
``util.regress`` (normal equations with a fixed 1e-6 ridge, util/util_state_space.py:33-61) or ``util.regress_alpha``
(Ridge, :64-94).  Parameter storage and the pickle payload keep the reference's dict layout."""

from __future__ import annotations

import pickle

import numpy as np
import torch

from .. import _capi as C
from ..engine import DeviceData, Engine

np_eps = np.finfo(float).eps


def check_status(status: int, alpha: float = 0.0):
    from ..framework.marginalizable_mixture_model import _raise_for_status

    _raise_for_status(status, alpha)


def unpack_row(row, d, l):
    """one packed parameter row [pi | m | S | A | Gamma | H | Lambda] -> the component's three dicts"""
    cut = np.cumsum([1, d, d * d, d * d, d * d, d * l, l * l])
    m, S, A, G, H, L = (np.array(p) for p in np.split(np.asarray(row, dtype=np.float64), cut)[1:7])
    return ({"mean": m, "cov": S.reshape(d, d)}, {"coeff": A.reshape(d, d), "covar": G.reshape(d, d)},
            {"coeff": H.reshape(d, l), "covar": L.reshape(l, l)})


class StateSpaceLinearGaussian:
    """Linear-Gaussian state-space model (linear dynamical system)."""

    def __init__(self, alpha: float = 0.0, *, device=None):
        self.alpha = alpha if alpha > 2 * np_eps else 0
        self.data = None
        self.state_init = self.state_model = self.measurement_model = None
        self._device = device

    def __str__(self):
        return "State space model with linear Gaussian components"

    # ---- parameters ------------------------------------------------------------------------------
    def _adopt(self, row, d, l):
        self.state_init, self.state_model, self.measurement_model = unpack_row(row, d, l)
        for holder, key, dim in ((self.state_init, "cov", d), (self.state_model, "covar", d),
                                 (self.measurement_model, "covar", l)):
            if dim == 1:  # np.cov of a single variable is 0-d in the reference
                holder[key] = holder[key].reshape(())
        return self

    def _packed(self, d, l):
        parts = (self.state_init["mean"], self.state_init["cov"], self.state_model["coeff"], self.state_model["covar"],
                 self.measurement_model["coeff"], self.measurement_model["covar"])
        return np.concatenate([[1.0]] + [np.ravel(np.asarray(p, dtype=np.float64)) for p in parts])[None]

    # ---- estimator interface -------------------------------------------------------------------------
    def fit(self, data):
        """all parameters from one pass over (states, measurements) on the device (SSLG:39-79)"""
        self.data = tuple(map(np.atleast_3d, data))
        eng = Engine.get_local(self._device)
        dd = DeviceData(*self.data, eng.device)
        one = torch.zeros(dd.n, dtype=torch.int32, device=eng.device)
        params, _, status = eng.mstep_ext(dd, one, 1, self.alpha)
        # an empty regression design is not an error for the component (util.regress solves the bare ridge)
        check_status(status & ~C.STATUS_EMPTY_FIT, self.alpha)
        return self._adopt(params[0].cpu().numpy(), dd.d, dd.l)

    def score(self, data=None):
        """log p(z_i, x_i) per trajectory, missing entries marginalised (SSLG:99-126)"""
        states, measurements = map(np.atleast_3d, self.data if data is None else data)
        eng = Engine.get_local(self._device)
        dd = DeviceData(states, measurements, eng.device)
        res = eng.estep(dd, self._packed(dd.d, dd.l), 1, dd.T, want_ll=True)
        check_status(res["status"], self.alpha)
        return res["ll"][0].cpu().numpy()

    # ---- persistence (payload keys of SSLG:81-97) --------------------------------------------------------
    def to_pickle(self) -> bytes:
        return pickle.dumps({"state_init": self.state_init, "state_model": self.state_model,
                             "measurement_model": self.measurement_model, "alpha": self.alpha})

    def from_pickle(self, p: bytes):
        payload = pickle.loads(p)
        for key in ("state_init", "state_model", "measurement_model"):
            setattr(self, key, payload[key])
        self.alpha = payload.get("alpha", 0)
        return self
