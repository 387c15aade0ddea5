"""Model-selection sweep over the number of clusters: the loop of the reference's
``model-selection.py`` (lines 45-72) on top of the batched-restart kernel path.

For every n_c the k-means start and ``n_starts`` random starts are fitted (the random ones in one
batched call, see ``mtm_sweep_fit``), the best model by expected complete-data log-likelihood is
kept, and its AIC / BIC / Q / LL are reported."""

from __future__ import annotations

import numpy as np

from .framework.marginalizable_mixture_model import MMLinGaussSS_marginalizable


def select_n_clusters(states, observations, n_cluster_list=range(1, 8), *, n_starts=10, n_steps=100, alpha=0.0,
                      device=None):
    """Returns (table, models): table is a dict of arrays keyed 'n_clusters', 'aic', 'bic', 'Q', 'LL',
    'n_free_params'; models maps n_c to the selected fitted model."""
    rows, models = [], {}
    shared = None
    for n_c in n_cluster_list:
        proto = MMLinGaussSS_marginalizable(n_clusters=n_c, states=states, observations=observations, alpha=alpha,
                                            device=device, _shared_data=shared)
        shared = proto._train_data()  # upload once, reuse for every n_c
        best = proto.train_with_multiple_random_starts(n_starts=n_starts, n_steps=n_steps, use_cache=False)
        models[n_c] = best
        rows.append((n_c, best.aic(), best.bic(), best.e_complete_data_log_lik(), best.model_log_likelihood(),
                     best.n_free_params))
    arr = np.array(rows, dtype=float)
    table = {k: arr[:, i] for i, k in enumerate(("n_clusters", "aic", "bic", "Q", "LL", "n_free_params"))}
    return table, models
