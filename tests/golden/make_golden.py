#!/usr/bin/env python3
"""Generate golden input/output vectors by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on the
GPU box):

    python tests/golden/make_golden.py

The reference is imported from /root/reference with ``sys.modules`` stubs for the two
absent plotting/reporting packages (matplotlib, statsmodels — only used by plotting and
``print_tests``) and ``np.alltrue`` / ``np.infty`` shimmed.  ``use_cache=False`` everywhere because the
reference directory is read-only.  Output: ``tests/golden/*.npz`` (committed).
"""

from __future__ import annotations

import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MTM_REFERENCE", "/root/reference")


def import_reference():
    for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.axes", "statsmodels", "statsmodels.api"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib.pyplot"].rcParams = {}
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if not hasattr(np, "alltrue"):
        np.alltrue = np.all
    try:
        np.infty
    except AttributeError:
        np.infty = np.inf  # removed in numpy 2; framework_extended/state_space_model_mixture.py still uses it
    sys.path.insert(0, REF)
    from framework import marginalizable_mixture_model as mmm
    from framework import marginalizable_state_space_model as ssm
    from util import util_state_space as uss

    return ssm, mmm, uss


def rand_spd(rng, k, ridge=0.3):
    B = rng.normal(size=(k, k))
    return B @ B.T / k + ridge * np.eye(k)


def rand_cluster(rng, d, l):
    return dict(m=2.0 * rng.normal(size=d), S=rand_spd(rng, d), A=rng.normal(scale=0.5, size=(d, d)),
                G=rand_spd(rng, d), H=rng.normal(size=(d, l)), L=rand_spd(rng, l))


def sample_lgssm(rng, n, T, p):
    """z_1 ~ N(m,S), z_{t+1} = z_t A + N(0,G), x_t = z_t H + N(0,L): the generative model of
    SSM:8-16 / SSM:776-836 (own sampler: the reference's breaks for 1-d states)."""
    d, l = p["H"].shape
    cs, cg, cl = (np.linalg.cholesky(p[k]) for k in ("S", "G", "L"))
    z = np.empty((T, n, d))
    x = np.empty((T, n, l))
    z[0] = p["m"] + rng.normal(size=(n, d)) @ cs.T
    for t in range(T):
        if t:
            z[t] = z[t - 1] @ p["A"] + rng.normal(size=(n, d)) @ cg.T
        x[t] = z[t] @ p["H"] + rng.normal(size=(n, l)) @ cl.T
    return z, x


def sample_mixture(ssm, rng, n, T, d, l, n_c, p_missing, sep=0.3):
    """per-cluster sampling + element-wise missingness.  Cluster parameters are a common base
    model moved a fraction ``sep`` towards independent random models: fully independent
    clusters (sep=1) are so well separated that hard EM collapses a cluster within 2-3
    iterations and the reference then raises on the NaN covariance."""
    base = rand_cluster(rng, d, l)
    comps = []
    for _ in range(n_c):
        p = rand_cluster(rng, d, l)
        comps.append({k: base[k] + sep * (p[k] - base[k]) for k in base})
    lab = rng.integers(0, n_c, size=n)
    z = np.empty((T, n, d))
    x = np.empty((T, n, l))
    for c, p in enumerate(comps):
        idx = np.flatnonzero(lab == c)
        z[:, idx, :], x[:, idx, :] = sample_lgssm(rng, idx.size, T, p)
    z[rng.random(z.shape) < p_missing] = np.nan
    x[rng.random(x.shape) < p_missing] = np.nan
    return z, x, lab, comps


def pack_model(mdl):
    """(n_c, 1+d+3d^2+dl+l^2) row-major [pi, m, S, A, G, H, L] per cluster."""
    rows = []
    for c in range(mdl.n_clusters):
        rows.append(np.concatenate([
            [mdl.cluster_propensities[c]], np.ravel(mdl.init_state_means[c]),
            np.ravel(mdl.init_state_covs[c]), np.ravel(mdl.transition_matrices[c]),
            np.ravel(mdl.transition_covs[c]), np.ravel(mdl.measurement_matrices[c]),
            np.ravel(mdl.measurement_covs[c])]))
    return np.stack(rows)


# ---------------------------------------------------------------------------
# 1. state-space algebra + marginalising log-density (SSM)
# ---------------------------------------------------------------------------


def golden_ssm(ssm):
    out = {}
    shapes = [(3, 2, 3), (4, 2, 4), (8, 4, 4), (5, 5, 3), (4, 1, 1), (6, 3, 2)]
    out["shapes"] = np.array(shapes)
    for k, (T, d, l) in enumerate(shapes):
        rng = np.random.default_rng(1000 + k)
        p = rand_cluster(rng, d, l)
        n = 40
        z, x = sample_lgssm(rng, n, T, p)
        z[rng.random(z.shape) < 0.25] = np.nan
        x[rng.random(x.shape) < 0.25] = np.nan
        # edge rows: complete, all missing, +-inf treated as missing, hidden states all missing,
        # suffix-missing (short trajectory), only the first step present
        z[:, 0, :], x[:, 0, :] = np.nan_to_num(z[:, 0, :], nan=0.3), np.nan_to_num(x[:, 0, :], nan=-0.2)
        z[:, 1, :], x[:, 1, :] = np.nan, np.nan
        z[0, 2, 0], x[-1, 2, 0] = np.inf, -np.inf
        z[:, 3, :] = np.nan
        z[T // 2:, 4, :], x[T // 2:, 4, :] = np.nan, np.nan
        z[1:, 5, :], x[1:, 5, :] = np.nan, np.nan
        mean = ssm.mm(T, p["m"], p["A"], p["H"])
        cov = ssm.CC(T, p["S"], p["A"], p["G"], p["H"], p["L"])
        Y = np.hstack((*z, *x))
        ll = ssm.multivariate_normal_log_likelihood(Y, mean, cov, np.zeros(n))
        meanx = ssm.mmX(T, p["m"], p["A"], p["H"])
        covx = ssm.CXX(T, p["S"], p["A"], p["G"], p["H"], p["L"])
        llx = ssm.multivariate_normal_log_likelihood(np.hstack(x), meanx, covx, np.zeros(n))
        for name, val in dict(z=z, x=x, mean=mean, cov=cov, ll=ll, meanx=meanx, covx=covx, llx=llx, **p).items():
            out[f"s{k}_{name}"] = val
    np.savez_compressed(os.path.join(HERE, "ssm_loglik.npz"), **out)
    print("ssm_loglik.npz written")


# ---------------------------------------------------------------------------
# 2. EM traces (MMM)
# ---------------------------------------------------------------------------


def traced_train(mmm, n_clusters, z, x, seed, init, alpha, n_steps, init_assign=None):
    """Run reference ``train`` while recording every E/M step."""
    mdl = mmm.MMLinGaussSS_marginalizable(n_clusters=n_clusters, states=z, observations=x,
                                          random_seed=seed, init=init, alpha=alpha)
    if init_assign is not None:
        mdl.cluster_assignment = np.array(init_assign)
    tr = dict(assign0=np.array(mdl.cluster_assignment), assign=[], switches=[], params=[])
    e0, m0 = mdl.E_step, mdl.M_step

    def e_wrap():
        s = e0()
        tr["assign"].append(np.array(mdl.cluster_assignment))
        tr["switches"].append(s)
        return s

    def m_wrap():
        m0()
        tr["params"].append(pack_model(mdl))

    mdl.E_step, mdl.M_step = e_wrap, m_wrap
    mdl.train(n_steps=n_steps)
    mdl.E_step, mdl.M_step = e0, m0
    return mdl, tr


def store_trace(out, key, mdl, tr):
    out[f"{key}_assign0"] = tr["assign0"]
    out[f"{key}_assign_hist"] = np.array(tr["assign"]).reshape(len(tr["assign"]), -1).astype(np.int8)
    out[f"{key}_switches"] = np.array(tr["switches"], dtype=np.int64)
    out[f"{key}_params_hist"] = np.array(tr["params"])
    out[f"{key}_params"] = pack_model(mdl)
    out[f"{key}_assign"] = np.array(mdl.cluster_assignment)
    a, probs, prenorm = mdl.mle_cluster_assignment(return_probs=True, return_prenormalized_log_probs=True)
    out[f"{key}_mle_assign"] = a
    out[f"{key}_probs"] = probs
    out[f"{key}_prenorm"] = prenorm
    out[f"{key}_ll"] = np.stack([mdl.conditional_log_likelihoods(c) for c in range(mdl.n_clusters)])
    out[f"{key}_Q"] = mdl.e_complete_data_log_lik()
    out[f"{key}_LL"] = mdl.model_log_likelihood()
    out[f"{key}_aic"] = mdl.aic()
    out[f"{key}_bic"] = mdl.bic()
    out[f"{key}_nfree"] = mdl.n_free_params
    T = mdl.n_timesteps
    keep = slice(0, 150)  # first 150 trajectories are enough for the prefix (T0) variants
    out[f"{key}_ll_T0"] = np.stack([
        np.stack([mdl.conditional_log_likelihoods_first_T0_steps(c, t0)[keep] for c in range(mdl.n_clusters)])
        for t0 in range(1, T + 1)])
    try:  # the reference's observation-only route raises for 1-d states (0-d covariance, SSM:253)
        ao, po = mdl.observations_mle_cluster_assignment(return_probs=True)
    except TypeError:
        print(key, ": reference observation-only path unavailable (1-d)")
        return
    out[f"{key}_obs_assign"] = ao
    out[f"{key}_obs_probs"] = po
    out[f"{key}_obs_ll"] = np.stack([mdl.observed_conditional_log_likelihoods(c) for c in range(mdl.n_clusters)])
    out[f"{key}_obs_ll_T0"] = np.stack([
        np.stack([mdl.observed_condl_log_lik_first_T0_steps(c, t0)[keep] for c in range(mdl.n_clusters)])
        for t0 in range(1, T + 1)])


def golden_em(ssm, mmm):
    out = {}
    cases = [
        # key, n, T, d, l, n_c, missing, seed, alpha, n_steps, sep
        ("c2s", 1200, 8, 4, 4, 4, 0.2, 42, 0.0, 1000, 0.3),    # config-2 shape, small n
        ("c3s", 2000, 8, 4, 4, 8, 0.2, 42, 0.0, 1000, 0.3),    # config-3 shape, small n
        ("slow", 1000, 8, 4, 4, 3, 0.2, 3, 0.0, 1000, 0.15),   # heavily overlapping clusters
        ("adnish", 600, 4, 2, 4, 4, 0.3, 101, 1.0, 1000, 0.4),  # ADNI shape, ridge
        ("mmmtest", 400, 10, 2, 3, 2, 0.1, 42, 0.1, 1000, 0.5),  # MMM self-test shape
        ("oned", 400, 5, 1, 1, 2, 0.15, 7, 0.0, 1000, 0.6),     # 1-d path
        ("short", 600, 6, 3, 2, 3, 0.2, 5, 0.0, 3, 0.3),        # exhausts n_steps
    ]
    names = []
    for key, n, T, d, l, n_c, pm, seed, alpha, n_steps, sep in cases:
        rng = np.random.default_rng(sum(map(ord, key)) + 1234)
        z, x, lab, comps = sample_mixture(ssm, rng, n, T, d, l, n_c, pm, sep)
        # Keep only well-conditioned traces: once a cluster shrinks to a handful of members its
        # covariance estimates turn singular, det() changes sign through rounding alone and the
        # reference's own path is no longer reproducible by anything but the same LAPACK build.
        for seed in range(seed, seed + 50):
            try:
                mdl, tr = traced_train(mmm, n_c, z, x, seed, "random", alpha, n_steps)
            except BaseException as err:  # noqa: BLE001 - the reference raises LinAlgError & co.
                print(key, "seed", seed, "raised", type(err).__name__)
                continue
            sizes = [np.bincount(a, minlength=n_c).min() for a in [tr["assign0"]] + tr["assign"]]
            if min(sizes) >= (0 if d == 1 else 20):  # scalar covariances cannot turn singular
                break
            print(key, "seed", seed, "rejected: min cluster size", min(sizes))
        else:
            raise SystemExit(f"no healthy trace for {key}")
        out[f"{key}_z"], out[f"{key}_x"] = z, x
        out[f"{key}_meta"] = np.array([n, T, d, l, n_c, seed, n_steps], dtype=np.int64)
        out[f"{key}_alpha"] = alpha
        store_trace(out, key, mdl, tr)
        names.append(key)
        print(key, "E-steps:", len(tr["switches"]), "switches:", tr["switches"][:8], "...")
        # held-out style predictions on a perturbed copy (fewer time steps + cognitive-only)
        zt, xt = z[: max(1, T - 1), : n // 3].copy(), x[: max(1, T - 1), : n // 3].copy()
        a, pr, pn = mdl.mle_cluster_assignment(states=zt, observations=xt, return_probs=True,
                                               return_prenormalized_log_probs=True)
        out[f"{key}_ho_assign"], out[f"{key}_ho_probs"], out[f"{key}_ho_prenorm"] = a, pr, pn
        a2, pr2 = mdl.mle_cluster_assignment(states=np.nan * np.ones_like(zt), observations=xt, return_probs=True)
        out[f"{key}_ho_nohid_assign"], out[f"{key}_ho_nohid_probs"] = a2, pr2
    # a fit that ends through the near-empty-cluster guard in mid-run (MMM:1346-1356):
    # small n, searched over seeds; the parameters stay those of the previous M-step
    rng = np.random.default_rng(4242)
    z, x, lab, comps = sample_mixture(ssm, rng, 160, 4, 2, 2, 4, 0.1, 1.0)
    for seed in range(200):
        try:
            mdl, tr = traced_train(mmm, 4, z, x, seed, "random", 0.0, 50)
        except BaseException:
            continue
        sw = tr["switches"]
        if sw and sw[-1] != 0 and len(sw) < 50 and len(tr["params"]) == len(sw) \
                and np.isfinite(pack_model(mdl)).all():
            key = "guardexit"
            out[f"{key}_z"], out[f"{key}_x"] = z, x
            out[f"{key}_meta"] = np.array([160, 4, 2, 2, 4, seed, 50], dtype=np.int64)
            out[f"{key}_alpha"] = 0.0
            for k2, v in dict(assign0=tr["assign0"], assign_hist=np.array(tr["assign"]).astype(np.int8),
                              switches=np.array(sw, dtype=np.int64), params_hist=np.array(tr["params"]),
                              params=pack_model(mdl), assign=np.array(mdl.cluster_assignment)).items():
                out[f"{key}_{k2}"] = v
            print("guardexit: seed", seed, "switches", sw, "counts", np.bincount(mdl.cluster_assignment, minlength=4))
            break
    else:
        raise SystemExit("no guard-exit case found")
    out["cases"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "em_traces.npz"), **out)
    print("em_traces.npz written")


# ---------------------------------------------------------------------------
# 3. multi-restart winner selection
# ---------------------------------------------------------------------------


def golden_restarts(ssm, mmm):
    """``train_with_multiple_random_starts`` (MMM:1368-1471) on two data sets:
    A: element-wise missingness.  The k-means start groups the trajectories whose z_1 is not
       fully finite into one cluster -> empty initial-state sample -> NaN mean -> every
       coordinate of that cluster counts as missing (isfinite(y - mu), SSM:769-770) ->
       ll = -0.0 -> the cluster swallows everything, the guard fires, and this degenerate
       model WINS the restart comparison with Q = n log(pi_c).  Reference behaviour to match.
    B: suffix-missing (ADNI-style lengths), z_1 always observed: a regular selection."""
    out = {}
    n_starts = 6
    for tag in ("A", "B"):
        rng = np.random.default_rng(77 if tag == "A" else 78)
        n, T, d, l, n_c = 600, 6, 3, 3, 3
        z, x, lab, comps = sample_mixture(ssm, rng, n, T, d, l, n_c, 0.2 if tag == "A" else 0.0, 0.4)
        if tag == "B":
            length = rng.integers(2, T + 1, size=n)
            for i in range(n):
                z[length[i]:, i, :] = np.nan
                x[length[i]:, i, :] = np.nan
            x[rng.random(x.shape) < 0.1] = np.nan
        proto = mmm.MMLinGaussSS_marginalizable(n_clusters=n_c, states=z, observations=x, alpha=0.0)
        best, obj = proto.train_with_multiple_random_starts(n_starts=n_starts, n_steps=100, use_cache=False,
                                                            return_objectives=True)
        km = mmm.MMLinGaussSS_marginalizable(n_clusters=n_c, states=z, observations=x, random_seed=0,
                                             init="kmeans")
        for k2, v in dict(z=z, x=x, meta=np.array([n, T, d, l, n_c, n_starts]), objectives=obj,
                          kmeans_assign0=np.array(km.cluster_assignment), best_params=pack_model(best),
                          best_assign=np.array(best.cluster_assignment), best_seed=best.random_seed,
                          best_Q=best.e_complete_data_log_lik(), best_LL=best.model_log_likelihood(),
                          best_aic=best.aic(), best_bic=best.bic()).items():
            out[f"{tag}_{k2}"] = v
        print("restarts", tag, "objectives", obj, "winner seed", best.random_seed)
    np.savez_compressed(os.path.join(HERE, "restarts.npz"), **out)
    print("restarts.npz written")


# ---------------------------------------------------------------------------
# 4. ADNI (config 1): one fold of inference-adni-xval.py, reduced n_starts
# ---------------------------------------------------------------------------


def golden_adni(ssm, mmm, uss):
    import sklearn.model_selection as skl_mdl_sel

    out = {}
    # same arrays as util_adni.get_trajectories() (util_adni.py:363-395); loaded directly because
    # util_adni needs a real matplotlib at import time
    arrs = np.load(os.path.join(REF, "data", "adni-trajectories.npz"), allow_pickle=True)
    z, x, ids = arrs["z"], arrs["x"], arrs["ids"]
    z_dim = z.shape[-1]
    markers = np.concatenate([z, x], axis=-1)
    n_splits, n_clusters, alpha, n_starts = 10, 4, 1.0, 8
    folds = list(skl_mdl_sel.KFold(n_splits=n_splits, shuffle=True, random_state=42).split(ids[0].ravel()))
    train_mask, test_mask = folds[0]
    mtrain, mtest = markers[:, train_mask].copy(), markers[:, test_mask].copy()
    mtrain[:, :, :z_dim], std = uss.standardize(mtrain[:, :, :z_dim], return_params=True)
    mtest[:, :, :z_dim] = uss.standardize(mtest[:, :, :z_dim], params=std)
    ztr, xtr = mtrain[:, :, :z_dim], mtrain[:, :, z_dim:]
    zte, xte = mtest[:, :, :z_dim], mtest[:, :, z_dim:]
    proto = mmm.MMLinGaussSS_marginalizable(n_clusters=n_clusters, states=ztr, observations=xtr,
                                            init="k-means", alpha=alpha)
    best, obj = proto.train_with_multiple_random_starts(n_starts=n_starts, use_cache=False,
                                                        return_objectives=True)
    km = mmm.MMLinGaussSS_marginalizable(n_clusters=n_clusters, states=ztr, observations=xtr,
                                         random_seed=0, init="kmeans", alpha=alpha)
    out.update(ztr=ztr, xtr=xtr, zte=zte, xte=xte, objectives=obj, alpha=alpha, n_starts=n_starts,
               kmeans_assign0=np.array(km.cluster_assignment), hex_hash=proto.hex_hash,
               best_params=pack_model(best), best_assign=np.array(best.cluster_assignment),
               best_seed=best.random_seed)
    # traced single fit from the k-means start (per-iteration parity)
    mdl, tr = traced_train(mmm, n_clusters, ztr, xtr, 0, "kmeans", alpha, 100)
    store_trace(out, "km", mdl, tr)
    # the five held-out maskings of inference-adni-xval.py:146-247
    a, pr, pn = best.mle_cluster_assignment(states=zte, observations=xte, return_probs=True,
                                            return_prenormalized_log_probs=True)
    out.update(te_assign=a, te_probs=pr, te_prenorm=pn)
    a0, p0 = best.mle_cluster_assignment(states=uss.mask_all_but_time_i(zte, 0),
                                         observations=uss.mask_all_but_time_i(xte, 0), return_probs=True)
    out.update(te0_assign=a0, te0_probs=p0)
    zp, xp = uss.parcellate_arrays(zte, xte)
    out.update(te_snap_assign=best.mle_cluster_assignment(states=zp, observations=xp),
               te_snap_index=best.cluster_assignment_index(states=zp, observations=xp),
               te_snapx_assign=best.mle_cluster_assignment(states=np.nan * np.ones_like(zp), observations=xp))
    znh = np.nan * np.ones_like(zte)
    out.update(te_nohid_assign=best.mle_cluster_assignment(states=znh, observations=xte),
               te_nohid_index=best.cluster_assignment_index(states=znh, observations=xte),
               te_obs_assign=best.observations_mle_cluster_assignment(observations=xte))
    np.savez_compressed(os.path.join(HERE, "adni_fold0.npz"), **out)
    print("adni_fold0.npz written; objectives", obj)


# ---------------------------------------------------------------------------
# 5. framework_extended linear-Gaussian component
# ---------------------------------------------------------------------------


def golden_lg_component(ssm):
    from framework_extended import state_space_model_linear_gaussian as ssm_lg

    out = {}
    rng = np.random.default_rng(2024)
    for tag, alpha in (("a0", 0.0), ("a1", 0.5)):
        z, x, lab, comps = sample_mixture(ssm, rng, 500, 6, 3, 4, 1, 0.15, 1.0)
        mdl = ssm_lg.StateSpaceLinearGaussian(alpha=alpha).fit((z, x))
        zt, xt = z[:, :200].copy(), x[:, :200].copy()
        zt[rng.random(zt.shape) < 0.2] = np.nan
        for k, v in dict(z=z, x=x, alpha=alpha, mean=mdl.state_init["mean"], cov=mdl.state_init["cov"],
                         A=mdl.state_model["coeff"], G=mdl.state_model["covar"], H=mdl.measurement_model["coeff"],
                         L=mdl.measurement_model["covar"], score=mdl.score(), zt=zt, xt=xt,
                         score_t=mdl.score((zt, xt))).items():
            out[f"{tag}_{k}"] = v
    np.savez_compressed(os.path.join(HERE, "lg_component.npz"), **out)
    print("lg_component.npz written")


# ---------------------------------------------------------------------------
# 6. secondary predictors built on the E-step
# ---------------------------------------------------------------------------


def golden_predictors(ssm, mmm):
    """cluster_propensities_over_time (MMM:532-577), observed_cluster_propensities_over_time (:1077-1123),
    initial_full_data_cluster_assignment (:909-941), predictions_from_initial_data (:943-995),
    one_step_ahead_predictions[_no_history] (:808-907), the T0-prefix likelihoods (:439-494, 997-1048) and
    aic / bic (:668-722) of one trained model, on its training set and on held-out trajectories."""
    out = {}
    rng = np.random.default_rng(77)
    n_tr, n_te, T, d, l, n_c = 400, 150, 5, 2, 3, 3
    z, x, lab, comps = sample_mixture(ssm, rng, n_tr + n_te, T, d, l, n_c, 0.2)
    z[0] = np.where(np.isnan(z[0]), 0.1, z[0])  # predictions_from_initial_data asserts on a complete first state
    ztr, xtr, zte, xte = z[:, :n_tr].copy(), x[:, :n_tr].copy(), z[:, n_tr:].copy(), x[:, n_tr:].copy()
    mdl = mmm.MMLinGaussSS_marginalizable(n_clusters=n_c, states=ztr, observations=xtr, random_seed=5)
    mdl.train(n_steps=40)
    assert min(np.bincount(mdl.cluster_assignment, minlength=n_c)) >= 20
    out.update(ztr=ztr, xtr=xtr, zte=zte, xte=xte, params=pack_model(mdl), assign=np.array(mdl.cluster_assignment),
               seed=5)
    out["pot_tr"] = mdl.cluster_propensities_over_time(states=ztr, observations=xtr)
    out["pot_te"] = mdl.cluster_propensities_over_time(states=zte, observations=xte)
    out["opot_tr"] = mdl.observed_cluster_propensities_over_time()
    out["opot_te"] = mdl.observed_cluster_propensities_over_time(observations=xte)
    out["init_tr"] = mdl.initial_full_data_cluster_assignment()
    out["init_te"] = mdl.initial_full_data_cluster_assignment(states=zte, observations=xte)
    ps, po = mdl.predictions_from_initial_data()
    out.update(pfi_s=ps, pfi_o=po)
    s1, o1 = mdl.one_step_ahead_predictions(states=zte, observations=xte)
    s2, o2 = mdl.one_step_ahead_predictions_no_history(states=zte, observations=xte)
    out.update(osa_s=s1, osa_o=o1, osanh_s=s2, osanh_o=o2)
    for T0 in (1, 3, T):
        out[f"ll_T{T0}"] = np.stack([mdl.conditional_log_likelihoods_first_T0_steps(c, T0) for c in range(n_c)])
        out[f"oll_T{T0}_te"] = np.stack([mdl.observed_condl_log_lik_first_T0_steps(c, T0, observations=xte)
                                         for c in range(n_c)])
    out.update(aic=mdl.aic(), bic=mdl.bic(), aic_te=mdl.aic(states=zte, observations=xte),
               bic_te=mdl.bic(states=zte, observations=xte), n_free=mdl.n_free_params,
               Q=mdl.e_complete_data_log_lik(), LL=mdl.model_log_likelihood())
    np.savez_compressed(os.path.join(HERE, "predictors.npz"), **out)
    print("predictors.npz written")


# ---------------------------------------------------------------------------
# 7. framework_extended mixture and classifier with linear-Gaussian components
# ---------------------------------------------------------------------------


def golden_extended(ssm):
    from framework_extended import state_space_model_classifier as ssm_clf
    from framework_extended import state_space_model_linear_gaussian as ssm_lg
    from framework_extended import state_space_model_mixture as ssm_mix

    import tempfile

    # the reference caches into <package>/tmp, which is read-only here: restarts (fit(use_cache=True) inside the
    # restart loop, :286-297) would raise in to_pickle and be silently dropped.  Point it at a scratch directory.
    ssm_mix.home_dir = tempfile.mkdtemp(prefix="mtm_ref_cache_")
    out = {}
    rng = np.random.default_rng(909)
    n_tr, n_te, T, d, l, n_c = 300, 120, 6, 2, 3, 2
    z, x, lab, comps = sample_mixture(ssm, rng, n_tr + n_te, T, d, l, n_c, 0.15, 1.0)
    ztr, xtr, zte, xte = z[:, :n_tr].copy(), x[:, :n_tr].copy(), z[:, n_tr:].copy(), x[:, n_tr:].copy()
    out.update(ztr=ztr, xtr=xtr, zte=zte, xte=xte, lab_tr=lab[:n_tr], lab_te=lab[n_tr:])
    for tag, kw in (("rand", dict(init="random", n_iter=50, n_restarts=2)), ("km", dict(init="k-means-all", n_iter=50))):
        zz = ztr if tag == "rand" else np.where(np.isnan(ztr), 0.0, ztr)  # k-means needs finite states
        mdl = ssm_mix.StateSpaceMixtureModel(n_clusters=n_c, data=(zz, xtr),
                                             component_model=ssm_lg.StateSpaceLinearGaussian,
                                             rng=np.random.default_rng(3)).fit(use_cache=False, **kw)
        pp, pn = mdl.predict_proba(data=(zte, xte), return_prenormalized_log_probs=True)
        out.update({f"{tag}_assign": np.array(mdl.cluster_assignment), f"{tag}_pi": np.array(mdl.cluster_propensities),
                    f"{tag}_score": mdl.score(), f"{tag}_mll": mdl.model_log_likelihood(),
                    f"{tag}_score_te": mdl.score(data=(zte, xte)), f"{tag}_proba_tr": mdl.predict_proba(),
                    f"{tag}_proba_te": pp, f"{tag}_prenorm_te": pn,
                    f"{tag}_pred_te": mdl.predict(data=(zte, xte), letters=False),
                    f"{tag}_index_te": mdl.cluster_assignment_index(cluster="B", data=(zte, xte)),
                    f"{tag}_A0": mdl.cluster_models[0].state_model["coeff"]})
    clf = ssm_clf.StateSpaceModelClassifier(component_model=ssm_lg.StateSpaceLinearGaussian).fit(
        data=(ztr, xtr), labels=lab[:n_tr])
    out.update(clf_score=clf.score(), clf_score_te=clf.score(data=(zte, xte)), clf_proba_te=clf.predict_proba((zte, xte)),
               clf_pred_te=clf.predict((zte, xte)), clf_pred_tr=clf.predict())
    np.savez_compressed(os.path.join(HERE, "extended.npz"), **out)
    print("extended.npz written; scores", out["rand_score"], out["km_score"], out["clf_score"])


# ---------------------------------------------------------------------------
# 7. conditioning: data far from the origin, rank-deficient regression designs
# ---------------------------------------------------------------------------


def golden_conditioning(ssm, mmm):
    """(a) "offset": a config-2-shaped fit on data moved far from the origin (z + 1e4, x - 3e3): one-pass raw
    moments would lose ~8 digits here; the reference works on the rows (lstsq, np.cov) and does not.
    (b) "dupcol": M_step on data whose third state coordinate duplicates the first: both regressions
    (MMM:1272-1308) have a rank-deficient design and np.linalg.lstsq(rcond=-1) (MMM:1194) returns the
    minimum-norm solution.  Only the M-step is stored: the fitted covariances are singular by construction,
    a following E-step is not meaningful.
    (c) static regress (MMM:1167-1200) on a duplicated-column design and on an underdetermined one."""
    out = {}
    rng = np.random.default_rng(20260)
    n, T, d, l, n_c = 1200, 8, 4, 4, 4
    z, x, lab, comps = sample_mixture(ssm, rng, n, T, d, l, n_c, 0.2, 0.3)
    z, x = z + 1.0e4, x - 3.0e3
    for seed in range(42, 92):
        try:
            mdl, tr = traced_train(mmm, n_c, z, x, seed, "random", 0.0, 1000)
        except BaseException as err:  # noqa: BLE001
            print("offset seed", seed, "raised", type(err).__name__)
            continue
        sizes = [np.bincount(a, minlength=n_c).min() for a in [tr["assign0"]] + tr["assign"]]
        if min(sizes) >= 20:
            break
    else:
        raise SystemExit("no healthy offset trace")
    key = "offset"
    out[f"{key}_z"], out[f"{key}_x"] = z, x
    out[f"{key}_meta"] = np.array([n, T, d, l, n_c, seed, 1000], dtype=np.int64)
    out[f"{key}_alpha"] = 0.0
    store_trace(out, key, mdl, tr)
    print("offset: seed", seed, "E-steps", len(tr["switches"]), tr["switches"][:8])

    rng = np.random.default_rng(20261)
    n, T, d, l, n_c = 500, 5, 3, 2, 2
    z, x, lab, comps = sample_mixture(ssm, rng, n, T, d, l, n_c, 0.15, 0.5)
    z[:, :, 2] = z[:, :, 0]
    mdl = mmm.MMLinGaussSS_marginalizable(n_clusters=n_c, states=z, observations=x, random_seed=5, alpha=0.0)
    mdl.M_step()
    out["dupcol_z"], out["dupcol_x"] = z, x
    out["dupcol_assign0"] = np.array(mdl.cluster_assignment)
    out["dupcol_params"] = pack_model(mdl)
    print("dupcol: A[0] =", np.round(mdl.transition_matrices[0], 4).tolist())

    rng = np.random.default_rng(20262)
    X = rng.normal(size=(300, 5))
    X[:, 3] = X[:, 1]
    X[:, 4] = 2.0 * X[:, 0] - X[:, 2]
    Y = X[:, :3] @ rng.normal(size=(3, 4)) + 0.2 * rng.normal(size=(300, 4))
    B, C = mmm.MMLinGaussSS_marginalizable.regress(X, Y)
    out["reg_dup_X"], out["reg_dup_Y"], out["reg_dup_B"], out["reg_dup_C"] = X, Y, B, C
    X2, Y2 = rng.normal(size=(3, 4)), rng.normal(size=(3, 2))
    B2, C2 = mmm.MMLinGaussSS_marginalizable.regress(X2, Y2)
    out["reg_under_X"], out["reg_under_Y"], out["reg_under_B"], out["reg_under_C"] = X2, Y2, B2, C2
    print("regress: rank-deficient |B| =", float(np.abs(B).max()), " underdetermined residual cov =", np.abs(C2).max())
    np.savez_compressed(os.path.join(HERE, "conditioning.npz"), **out)
    print("conditioning.npz written")


# ---------------------------------------------------------------------------
# 8. model selection (model-selection.py:45-72)
# ---------------------------------------------------------------------------


def golden_model_selection(ssm, mmm):
    """The loop of model-selection.py: for n_c in a range, the best of a k-means start and a few random starts,
    then its AIC / BIC / Q / LL (MSEL:45-72), on a small 3-cluster data set with suffix-missing lengths."""
    rng = np.random.default_rng(31337)
    n, T, d, l = 900, 5, 2, 3
    z, x, lab, comps = sample_mixture(ssm, rng, n, T, d, l, 3, 0.0, 0.8)
    length = rng.integers(2, T + 1, size=n)
    for i in range(n):
        z[length[i]:, i, :] = np.nan
        x[length[i]:, i, :] = np.nan
    x[rng.random(x.shape) < 0.1] = np.nan
    rows, n_starts = [], 4
    out = dict(z=z, x=x, n_starts=n_starts)
    for n_c in range(1, 5):
        best = mmm.MMLinGaussSS_marginalizable(n_clusters=n_c, states=z, observations=x).train_with_multiple_random_starts(
            n_starts=n_starts, n_steps=100, use_cache=False)
        rows.append([n_c, best.aic(), best.bic(), best.e_complete_data_log_lik(), best.model_log_likelihood(),
                     best.n_free_params, best.random_seed])
        out[f"assign_{n_c}"] = np.array(best.cluster_assignment)
        print("model selection n_c", n_c, rows[-1])
    out["table"] = np.array(rows, dtype=float)
    np.savez_compressed(os.path.join(HERE, "model_selection.npz"), **out)
    print("model_selection.npz written")


if __name__ == "__main__":
    ssm, mmm, uss = import_reference()
    if "msel" in sys.argv[1:]:
        golden_model_selection(ssm, mmm)
        sys.exit(0)
    if "cond" in sys.argv[1:]:
        golden_conditioning(ssm, mmm)
        sys.exit(0)
    if "ext" in (sys.argv[1:] or ["ext"]):
        golden_extended(ssm)
    which = sys.argv[1:] or ["ssm", "em", "restarts", "adni", "lg", "pred"]
    if "pred" in which:
        golden_predictors(ssm, mmm)
    if "ssm" in which:
        golden_ssm(ssm)
    if "em" in which:
        golden_em(ssm, mmm)
    if "restarts" in which:
        golden_restarts(ssm, mmm)
    if "adni" in which:
        golden_adni(ssm, mmm, uss)
    if "lg" in which:
        golden_lg_component(ssm)
