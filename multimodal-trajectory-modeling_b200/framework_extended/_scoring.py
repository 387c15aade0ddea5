"""(n_models, n) conditional log-likelihood matrix of a list of component models on one data set.

When every component is this package's ``StateSpaceLinearGaussian`` the whole matrix comes from ONE E-step launch
(each component = one parameter row with pi = 1) instead of one ``score`` call — one upload of the data and one
launch — per component, as ``StateSpaceMixtureModel._E_step`` / ``predict_proba`` / ``score`` of the reference do
(framework_extended/state_space_model_mixture.py:110-132, 406-474).  Any other component type is asked for its
own ``score``."""

from __future__ import annotations

import numpy as np

from ..engine import DeviceData, Engine
from .state_space_model_linear_gaussian import StateSpaceLinearGaussian, check_status


def all_linear_gaussian(models) -> bool:
    return bool(models) and all(type(m) is StateSpaceLinearGaussian for m in models)


def score_matrix(models, data, device=None) -> np.ndarray:
    if all_linear_gaussian(models):
        eng = Engine.get_local(device)
        dd = data if isinstance(data, DeviceData) else DeviceData(*map(np.atleast_3d, data), eng.device)
        packed = np.concatenate([m._packed(dd.d, dd.l) for m in models])
        res = eng.estep(dd, packed, len(models), dd.T, want_ll=True)
        check_status(res["status"])  # every failure bit, like the mixture class
        return dd.from_resident(res["ll"]).cpu().numpy()
    if isinstance(data, DeviceData):
        data = (data.from_resident(data.z).cpu().numpy(), data.from_resident(data.x).cpu().numpy())
    states, measurements = map(np.atleast_3d, data)
    return np.stack([np.asarray(m.score((states, measurements))) for m in models])
