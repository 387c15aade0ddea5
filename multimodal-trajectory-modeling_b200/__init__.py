"""B200-native hard-assignment EM for mixtures of linear-Gaussian state-space models.

Drop-in for the EM / predict path of burkh4rt/Multimodal-Trajectory-Modeling
(``framework/marginalizable_mixture_model.py`` + ``framework/marginalizable_state_space_model.py``).

The directory name carries hyphens, so import it through ``importlib`` or the ``mtm_b200``
alias module at the repository root::

    import mtm_b200
    from mtm_b200.framework import marginalizable_mixture_model as mixmodel
    mdl = mixmodel.MMLinGaussSS_marginalizable(n_clusters=4, states=z, observations=x).train()
"""

from . import _capi  # noqa: F401
from ._capi import MtmError, build  # noqa: F401

__all__ = ["MtmError", "build", "Engine", "DeviceData", "MMLinGaussSS_marginalizable"]


def __getattr__(name):
    # torch is imported lazily: `import mtm_b200` must work for build()/symbol checks alone
    if name in ("Engine", "DeviceData"):
        from . import engine

        return getattr(engine, name)
    if name == "MMLinGaussSS_marginalizable":
        from .framework.marginalizable_mixture_model import MMLinGaussSS_marginalizable

        return MMLinGaussSS_marginalizable
    raise AttributeError(name)
