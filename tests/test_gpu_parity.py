"""Parity of the CUDA path (through the drop-in class and the C ABI) with the reference:
golden vectors produced by the unmodified reference (tests/golden) and the CPU oracle.

Bar (BASELINE.json north_star): cluster assignments, switch counts and iteration counts
bit-exact; log-likelihoods, parameters, probabilities and scores within 1e-9 relative."""

import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

RTOL = 1e-9


FLOOR = 1e-6


def relerr(a, b, floor=FLOOR):
    """max |a - b| / max(|b|, floor): TRUE relative error (north_star: 1e-9 relative); the explicit absolute floor
    only shields entries that are zero up to rounding (|b| < 1e-6 is compared at 1e-15 absolute)."""
    a, b = np.atleast_1d(np.asarray(a, dtype=float)), np.atleast_1d(np.asarray(b, dtype=float))
    assert a.shape == b.shape, (a.shape, b.shape)
    with np.errstate(all="ignore"):
        d = np.abs(a - b) / np.maximum(floor, np.abs(b))
    d[np.isnan(a) & np.isnan(b)] = 0.0
    d[a == b] = 0.0  # equal infinities
    return float(np.max(d)) if d.size else 0.0


@pytest.fixture(scope="module")
def mix():
    import mtm_b200  # noqa: F401
    from mtm_b200.framework import marginalizable_mixture_model as mixmodel

    return mixmodel


@pytest.fixture(scope="module")
def eng():
    import mtm_b200

    return mtm_b200.Engine.get()


def _case(g_em, key):
    n, T, d, l, n_c, seed, n_steps = (int(v) for v in g_em[f"{key}_meta"])
    return dict(z=g_em[f"{key}_z"], x=g_em[f"{key}_x"], n=n, T=T, d=d, l=l, n_c=n_c, seed=seed,
                n_steps=n_steps, alpha=float(g_em[f"{key}_alpha"]))


def _set_params(mdl, packed):
    mdl._unpack_params(np.asarray(packed))


def _packed_single(g, k):
    return np.concatenate([[1.0], g(f"s{k}_m").ravel(), g(f"s{k}_S").ravel(), g(f"s{k}_A").ravel(),
                           g(f"s{k}_G").ravel(), g(f"s{k}_H").ravel(), g(f"s{k}_L").ravel()])[None]


def test_loglik_golden(g_ssm, eng):
    """SSM-level vectors: every (T, d, l) shape incl. the generic-dimension kernel, NaN / +-inf /
    all-missing / suffix-missing rows, joint and observation-only."""
    import mtm_b200

    for k, (T, d, l) in enumerate(g_ssm["shapes"]):
        g = lambda name: g_ssm[name]
        P = _packed_single(g, k)
        data = mtm_b200.DeviceData(g(f"s{k}_z"), g(f"s{k}_x"), eng.device)
        res = eng.estep(data, P, 1, T, want_ll=True)
        assert res["status"] == 0
        ll = res["ll"].cpu().numpy()[0]
        assert relerr(ll, g(f"s{k}_ll")) < RTOL, (T, d, l)
        assert ll[1] == 0.0  # all-missing row -> (-)0.0 like the reference
        xo = mtm_b200.DeviceData(None, g(f"s{k}_x"), eng.device)
        llx = eng.estep(xo, P, 1, T, want_ll=True, d=int(d))["ll"].cpu().numpy()[0]
        assert relerr(llx, g(f"s{k}_llx")) < RTOL, (T, d, l)


@pytest.fixture(params=[1, 2], ids=["stats-auto", "stats-delta"])
def delta_mode(request, eng):
    """The golden fits run twice: as shipped (these sets are small, so every iteration re-sums every cluster) and
    with delta statistics forced on (mtm_set_delta_stats 2) — the reference's traces must come out either way."""
    eng.set_delta_stats(request.param)
    yield request.param
    eng.set_delta_stats(1)


@pytest.mark.parametrize("key", ["c2s", "c3s", "slow", "adnish", "mmmtest", "oned", "short", "guardexit"])
def test_em_trace(g_em, mix, key, delta_mode):
    c = _case(g_em, key)
    mdl = mix.MMLinGaussSS_marginalizable(n_clusters=c["n_c"], states=c["z"], observations=c["x"],
                                          random_seed=c["seed"], alpha=c["alpha"])
    assert np.array_equal(mdl.cluster_assignment, g_em[f"{key}_assign0"])
    mdl.train(n_steps=c["n_steps"])
    info = mdl.fit_info
    assert info["switches"] == list(g_em[f"{key}_switches"]), info
    assert info["n_e_steps"] == len(g_em[f"{key}_switches"])
    assert np.array_equal(mdl.cluster_assignment, g_em[f"{key}_assign"])
    assert mdl.cluster_assignment.dtype == np.int64
    assert relerr(mdl._packed_params(), g_em[f"{key}_params"]) < RTOL
    sw = list(g_em[f"{key}_switches"])
    expect_exit = "converged" if sw[-1] == 0 else ("exhausted" if len(sw) == c["n_steps"] else "guard")
    assert info["exit"] == expect_exit
    assert {"short": "exhausted", "guardexit": "guard", "c2s": "converged"}.get(key, expect_exit) == expect_exit
    if expect_exit == "guard":
        return  # parameters are those of the previous M-step; nothing else was stored
    # outputs of the trained model
    a, probs, prenorm = mdl.mle_cluster_assignment(return_probs=True, return_prenormalized_log_probs=True)
    assert np.array_equal(a, g_em[f"{key}_mle_assign"])
    assert relerr(probs, g_em[f"{key}_probs"]) < RTOL
    assert relerr(prenorm, g_em[f"{key}_prenorm"]) < RTOL
    assert relerr(mdl.log_likelihood_matrix(), g_em[f"{key}_ll"]) < RTOL
    assert relerr(mdl.conditional_log_likelihoods(0), g_em[f"{key}_ll"][0]) < RTOL
    assert relerr(mdl.e_complete_data_log_lik(), g_em[f"{key}_Q"]) < RTOL
    assert relerr(mdl.model_log_likelihood(), g_em[f"{key}_LL"]) < RTOL
    assert relerr(mdl.aic(), g_em[f"{key}_aic"]) < RTOL
    assert relerr(mdl.bic(), g_em[f"{key}_bic"]) < RTOL
    assert mdl.n_free_params == int(g_em[f"{key}_nfree"])
    T = c["T"]
    for t0 in range(1, T + 1):
        got = mdl.log_likelihood_matrix(states=c["z"][:, :150], observations=c["x"][:, :150], T0=t0)
        assert relerr(got, g_em[f"{key}_ll_T0"][t0 - 1]) < RTOL
    # held-out style calls: fewer time steps than training, then hidden states all missing
    zt, xt = c["z"][: max(1, T - 1), : c["n"] // 3], c["x"][: max(1, T - 1), : c["n"] // 3]
    a, pr, pn = mdl.mle_cluster_assignment(states=zt, observations=xt, return_probs=True,
                                           return_prenormalized_log_probs=True)
    assert np.array_equal(a, g_em[f"{key}_ho_assign"])
    assert relerr(pr, g_em[f"{key}_ho_probs"]) < RTOL
    assert relerr(pn, g_em[f"{key}_ho_prenorm"]) < RTOL
    a2, pr2 = mdl.mle_cluster_assignment(states=np.nan * np.ones_like(zt), observations=xt, return_probs=True)
    assert np.array_equal(a2, g_em[f"{key}_ho_nohid_assign"])
    assert relerr(pr2, g_em[f"{key}_ho_nohid_probs"]) < RTOL
    if f"{key}_obs_ll" in g_em.files:  # the reference's observation-only route (not for 1-d)
        ao, po = mdl.observations_mle_cluster_assignment(return_probs=True)
        assert np.array_equal(ao, g_em[f"{key}_obs_assign"])
        assert relerr(po, g_em[f"{key}_obs_probs"]) < RTOL
        for cidx in range(c["n_c"]):
            assert relerr(mdl.observed_conditional_log_likelihoods(cidx), g_em[f"{key}_obs_ll"][cidx]) < RTOL
        for t0 in (1, T // 2 + 1, T):
            got = mdl.observed_condl_log_lik_first_T0_steps(1, t0, observations=c["x"][:, :150])
            assert relerr(got, g_em[f"{key}_obs_ll_T0"][t0 - 1][1]) < RTOL


@pytest.mark.parametrize("key", ["c2s", "adnish"])
def test_single_steps(g_em, mix, key):
    """E_step() / M_step() called one at a time reproduce every iterate of the reference."""
    c = _case(g_em, key)
    hist = g_em[f"{key}_assign_hist"].astype(np.int64)
    ph = g_em[f"{key}_params_hist"]
    sw = g_em[f"{key}_switches"]
    mdl = mix.MMLinGaussSS_marginalizable(n_clusters=c["n_c"], states=c["z"], observations=c["x"],
                                          random_seed=c["seed"], alpha=c["alpha"])
    mdl.M_step()
    assert relerr(mdl._packed_params(), ph[0]) < RTOL
    for it in range(len(sw)):
        assert mdl.E_step() == int(sw[it])
        assert np.array_equal(mdl.cluster_assignment, hist[it])
        if it + 1 < len(ph):
            mdl.M_step()
            assert relerr(mdl._packed_params(), ph[it + 1]) < RTOL


@pytest.mark.parametrize("tag", ["A", "B"])
def test_multi_restart(g_restarts, mix, tag):
    """train_with_multiple_random_starts: objectives of every start, the winner (strict >, k-means
    incumbent) and — data set A — the NaN-mean-cluster quirk that makes the degenerate k-means
    start win in the reference."""
    z, x = g_restarts[f"{tag}_z"], g_restarts[f"{tag}_x"]
    n, T, d, l, n_c, n_starts = (int(v) for v in g_restarts[f"{tag}_meta"])
    proto = mix.MMLinGaussSS_marginalizable(n_clusters=n_c, states=z, observations=x)
    km = mix.MMLinGaussSS_marginalizable(n_clusters=n_c, states=z, observations=x, random_seed=0, init="kmeans")
    assert np.array_equal(km.cluster_assignment, g_restarts[f"{tag}_kmeans_assign0"])
    best, obj = proto.train_with_multiple_random_starts(n_starts=n_starts, n_steps=100, use_cache=False,
                                                        return_objectives=True)
    assert relerr(obj, g_restarts[f"{tag}_objectives"]) < RTOL
    assert best.random_seed == int(g_restarts[f"{tag}_best_seed"])
    assert np.array_equal(best.cluster_assignment, g_restarts[f"{tag}_best_assign"])
    assert relerr(best._packed_params(), g_restarts[f"{tag}_best_params"]) < RTOL
    assert relerr(best.e_complete_data_log_lik(), g_restarts[f"{tag}_best_Q"]) < RTOL
    assert relerr(best.model_log_likelihood(), g_restarts[f"{tag}_best_LL"]) < RTOL
    assert relerr(best.aic(), g_restarts[f"{tag}_best_aic"]) < RTOL
    assert relerr(best.bic(), g_restarts[f"{tag}_best_bic"]) < RTOL


def test_adni_fold(g_adni, mix):
    """config 1: ADNI fold 0 of inference-adni-xval.py (alpha = 1, k-means + 8 random starts) and
    its held-out maskings (XVAL:146-247)."""
    ztr, xtr, zte, xte = g_adni["ztr"], g_adni["xtr"], g_adni["zte"], g_adni["xte"]
    alpha = float(g_adni["alpha"])
    proto = mix.MMLinGaussSS_marginalizable(n_clusters=4, states=ztr, observations=xtr, init="k-means", alpha=alpha)
    assert proto.hex_hash == str(g_adni["hex_hash"])
    km = mix.MMLinGaussSS_marginalizable(n_clusters=4, states=ztr, observations=xtr, random_seed=0, init="kmeans",
                                         alpha=alpha)
    assert np.array_equal(km.cluster_assignment, g_adni["km_assign0"])
    km.train(n_steps=100)
    assert km.fit_info["switches"] == list(g_adni["km_switches"])
    assert np.array_equal(km.cluster_assignment, g_adni["km_assign"])
    assert relerr(km._packed_params(), g_adni["km_params"]) < RTOL
    best, obj = proto.train_with_multiple_random_starts(n_starts=int(g_adni["n_starts"]), use_cache=False,
                                                        return_objectives=True)
    assert relerr(obj, g_adni["objectives"]) < RTOL
    assert best.random_seed == int(g_adni["best_seed"])
    assert np.array_equal(best.cluster_assignment, g_adni["best_assign"])
    assert relerr(best._packed_params(), g_adni["best_params"]) < RTOL
    a, pr, pn = best.mle_cluster_assignment(states=zte, observations=xte, return_probs=True,
                                            return_prenormalized_log_probs=True)
    assert np.array_equal(a, g_adni["te_assign"])
    assert relerr(pr, g_adni["te_probs"]) < RTOL
    assert relerr(pn, g_adni["te_prenorm"]) < RTOL
    mask0 = lambda arr: np.concatenate([arr[:1], np.full_like(arr[1:], np.nan)], axis=0)
    a0, p0 = best.mle_cluster_assignment(states=mask0(zte), observations=mask0(xte), return_probs=True)
    assert np.array_equal(a0, g_adni["te0_assign"])
    assert relerr(p0, g_adni["te0_probs"]) < RTOL
    snap = lambda arr: np.concatenate([np.where(np.arange(arr.shape[0])[:, None, None] == i, arr, np.nan)
                                       for i in range(arr.shape[0])], axis=1)
    zp, xp = snap(zte), snap(xte)
    assert np.array_equal(best.mle_cluster_assignment(states=zp, observations=xp), g_adni["te_snap_assign"])
    assert relerr(best.cluster_assignment_index(states=zp, observations=xp), g_adni["te_snap_index"]) < RTOL
    assert np.array_equal(best.mle_cluster_assignment(states=np.nan * np.ones_like(zp), observations=xp),
                          g_adni["te_snapx_assign"])
    # the same snapshots without materialising the n T pseudo-trajectories (one single-step launch per visit)
    a_s, pr_s, pn_s = best.mle_cluster_assignment_snapshots(states=zte, observations=xte, return_probs=True,
                                                            return_prenormalized_log_probs=True)
    a_m, pr_m, pn_m = best.mle_cluster_assignment(states=zp, observations=xp, return_probs=True,
                                                  return_prenormalized_log_probs=True)
    assert np.array_equal(a_s, g_adni["te_snap_assign"]) and np.array_equal(a_s, a_m)
    assert relerr(pr_s, pr_m) < RTOL and relerr(pn_s, pn_m) < RTOL
    assert relerr(best.cluster_assignment_index_snapshots(states=zte, observations=xte), g_adni["te_snap_index"]) < RTOL
    assert np.array_equal(best.mle_cluster_assignment_snapshots(states=zte, observations=xte, hidden_states=False),
                          g_adni["te_snapx_assign"])
    znh = np.nan * np.ones_like(zte)
    assert np.array_equal(best.mle_cluster_assignment(states=znh, observations=xte), g_adni["te_nohid_assign"])
    assert relerr(best.cluster_assignment_index(states=znh, observations=xte), g_adni["te_nohid_index"]) < RTOL
    assert np.array_equal(best.observations_mle_cluster_assignment(observations=xte), g_adni["te_obs_assign"])


# -------------------------------------------------------------------------------------------------
# against the CPU oracle on fresh seeded inputs, and size-independent properties at scale
# -------------------------------------------------------------------------------------------------


def _synthetic(n, T, d, l, n_c, seed, p_missing=0.2, sep=0.3):
    rng = np.random.default_rng(seed)

    def spd(k):
        B = rng.normal(size=(k, k))
        return B @ B.T / k + 0.3 * np.eye(k)

    def comp():
        return dict(m=2 * rng.normal(size=d), S=spd(d), A=rng.normal(scale=0.5, size=(d, d)), G=spd(d),
                    H=rng.normal(size=(d, l)), L=spd(l))

    base = comp()
    comps = []
    for _ in range(n_c):
        p = comp()
        comps.append({k: base[k] + sep * (p[k] - base[k]) for k in base})
    lab = rng.integers(0, n_c, size=n)
    z, x = np.empty((T, n, d)), np.empty((T, n, l))
    for c, p in enumerate(comps):
        idx = np.flatnonzero(lab == c)
        cs, cg, cl = (np.linalg.cholesky(p[k]) for k in ("S", "G", "L"))
        zc = np.empty((T, idx.size, d))
        zc[0] = p["m"] + rng.normal(size=(idx.size, d)) @ cs.T
        for t in range(T):
            if t:
                zc[t] = zc[t - 1] @ p["A"] + rng.normal(size=(idx.size, d)) @ cg.T
            x[t, idx] = zc[t] @ p["H"] + rng.normal(size=(idx.size, l)) @ cl.T
        z[:, idx] = zc
    z[rng.random(z.shape) < p_missing] = np.nan
    x[rng.random(x.shape) < p_missing] = np.nan
    return z, x


@pytest.mark.parametrize("shape", [(8, 4, 4, 4, 30011), (4, 2, 4, 4, 7777), (6, 3, 5, 3, 4099), (3, 6, 2, 2, 2500)])
def test_fit_vs_oracle(mix, shape):
    """Whole EM fits against the oracle (numpy M-step + C E-step) on inputs the goldens do not
    cover: ragged n (not a multiple of the 128-trajectory tile), generic-dimension kernels."""
    from oracle import cport as oc
    from oracle import dense as od

    T, d, l, n_c, n = shape
    z, x = _synthetic(n, T, d, l, n_c, seed=n)
    a0 = od.random_init(n, n_c, 11)
    ref = od.train(z, x, a0, n_c, alpha=0.0, n_steps=6, e_step_fn=oc.e_step_seq)
    mdl = mix.MMLinGaussSS_marginalizable(n_clusters=n_c, states=z, observations=x, random_seed=11)
    mdl.train(n_steps=6)
    assert mdl.fit_info["switches"] == ref.switches
    assert np.array_equal(mdl.cluster_assignment, ref.assign)
    assert relerr(mdl._packed_params(), ref.params.packed()) < RTOL
    # dense (reference-formulation) check of the final log-likelihoods on a sample
    sub = slice(0, 400)
    lld = oc.dense_loglik_matrix(ref.params, z[:, sub], x[:, sub])
    assert relerr(mdl.log_likelihood_matrix(states=z[:, sub], observations=x[:, sub]), lld) < RTOL


def test_scale_properties(mix, eng):
    """config-2 size (1M trajectories, T=8, d=l=4, n_c=4): properties that do not need an oracle
    pass at full size, plus an oracle check on a strided sample."""
    import mtm_b200
    from oracle import cport as oc

    n, T, d, l, n_c = 1_000_000, 8, 4, 4, 4
    zs, xs = _synthetic(50_000, T, d, l, n_c, seed=5)
    reps = n // zs.shape[1]
    z = np.ascontiguousarray(np.tile(zs, (1, reps, 1)))
    x = np.ascontiguousarray(np.tile(xs, (1, reps, 1)))
    rng = np.random.default_rng(0)
    # break the exact periodicity of the tiling with fresh missingness
    z[rng.random(z.shape) < 0.02] = np.nan
    mdl = mix.MMLinGaussSS_marginalizable(n_clusters=n_c, states=z, observations=x, random_seed=42)
    mdl.train(n_steps=3)
    assert mdl.fit_info["n_e_steps"] == 3
    P = mdl._packed_params()
    data = mdl._train_data()  # resident (pattern) order; from_resident gives the caller's order back
    assert data.perm is not None
    res = eng.estep(data, P, n_c, T, want_ll=True)
    ll = data.from_resident(res["ll"]).cpu().numpy()
    a = data.from_resident(res["assign"]).cpu().numpy()
    # (a) oracle on a strided sample of the full-size run
    idx = np.arange(0, n, 97)[:8000]
    llo = oc.seq_loglik_matrix(P, z[:, idx], x[:, idx])
    assert relerr(ll[:, idx], llo) < RTOL
    assert np.array_equal(a[idx], oc.assign(llo, P[:, 0]))
    # (b) permutation equivariance: shuffling trajectories shuffles the outputs — bit for bit, although the
    # pattern-ordered resident set takes the warp-uniform point-mass shortcuts and the shuffled one does not
    perm = rng.permutation(n)
    dp = mtm_b200.DeviceData(z[:, perm], x[:, perm], eng.device)
    rp = eng.estep(dp, P, n_c, T, want_ll=True)
    assert np.array_equal(rp["ll"].cpu().numpy(), ll[:, perm])
    assert np.array_equal(rp["assign"].cpu().numpy(), a[perm])
    # (c) switch count = number of changed labels; cluster sizes consistent
    old_h = rng.integers(0, n_c, n).astype(np.int32)
    old = data.to_resident(old_h)
    r2 = eng.estep(data, P, n_c, T, assign=old, count_switches=True)
    assert np.array_equal(data.from_resident(old).cpu().numpy(), a)
    assert r2["n_switches"] == int(np.sum(old_h != a))
    assert r2["min_count"] == int(np.bincount(a, minlength=n_c).min())
    # (d) hidden states all NaN through the joint kernel == observation-only kernel
    zn = mtm_b200.DeviceData(np.full((T, 4096, d), np.nan), x[:, :4096], eng.device)
    xo = mtm_b200.DeviceData(None, x[:, :4096], eng.device)
    l1 = eng.estep(zn, P, n_c, T, want_ll=True)["ll"].cpu().numpy()
    l2 = eng.estep(xo, P, n_c, T, want_ll=True, d=d)["ll"].cpu().numpy()
    assert relerr(l1, l2) < 1e-13
    # (e) statistics are additive over shards (what the multi-GPU all-reduce relies on)
    at = torch.from_numpy(a.astype(np.int32)).to(eng.device)
    at_res = data.to_resident(a.astype(np.int32))
    full = eng.mstats(data, at_res, n_c).cpu().numpy()
    h = n // 2 + 13
    d1 = mtm_b200.DeviceData(z[:, :h], x[:, :h], eng.device)
    d2 = mtm_b200.DeviceData(z[:, h:], x[:, h:], eng.device)
    s1 = eng.mstats(d1, at[:h].contiguous(), n_c).cpu().numpy()
    s2 = eng.mstats(d2, at[h:].contiguous(), n_c).cpu().numpy()
    assert relerr(s1 + s2, full) < 1e-12
    # counts inside the statistics are exact
    assert np.array_equal(full[:, 0], np.bincount(a, minlength=n_c).astype(float))
    # (f) M-step from the statistics == numpy M-step of the oracle on a sample-sized problem
    pf, _ = eng.finalize(torch.from_numpy(full).to(eng.device), n_c, d, l, n, 0.0)
    p2, _, st = eng.mstep(data, at_res, n_c, 0.0)
    # (mtm_mstats returns plain moments; mtm_mstep takes them about the pivot: same parameters up to rounding)
    assert st == 0 and relerr(pf.cpu().numpy(), p2.cpu().numpy()) < 1e-12


def test_edge_inputs(mix, eng):
    import mtm_b200
    from oracle import cport as oc

    z, x = _synthetic(300, 5, 4, 4, 3, seed=3)
    mdl = mix.MMLinGaussSS_marginalizable(n_clusters=3, states=z, observations=x, random_seed=1)
    mdl.train(n_steps=4)
    P = mdl._packed_params()
    for n in (1, 31, 127, 128, 129, 300):
        for T0 in (1, 5):
            got = mdl.log_likelihood_matrix(states=z[:, :n], observations=x[:, :n], T0=T0)
            assert relerr(got, oc.seq_loglik_matrix(P, z[:, :n], x[:, :n], T0=T0)) < RTOL
    # a single time step (T = 1) and +-inf entries
    z1, x1 = z[:1].copy(), x[:1].copy()
    z1[0, 0, 0], x1[0, 1, 2] = np.inf, -np.inf
    got = mdl.log_likelihood_matrix(states=z1, observations=x1, T0=1)
    assert relerr(got, oc.seq_loglik_matrix(P, z1, x1)) < RTOL
    # near-empty cluster before the first M-step: nothing runs (MMM:1328-1336)
    m2 = mix.MMLinGaussSS_marginalizable(n_clusters=3, states=z, observations=x, random_seed=1)
    m2.cluster_assignment = np.r_[np.zeros(150, int), np.ones(147, int), 2 * np.ones(3, int)]
    before = m2._packed_params().copy()
    m2.train()
    assert m2.fit_info["exit"] == "guard0" and m2.last_trained is None
    assert np.array_equal(m2._packed_params(), before)
    # n_steps = 0: guard + one M-step, no E-step
    m3 = mix.MMLinGaussSS_marginalizable(n_clusters=3, states=z, observations=x, random_seed=1)
    m3.train(n_steps=0)
    assert m3.fit_info["n_e_steps"] == 0 and m3.fit_info["n_m_steps"] == 1
    # non-finite covariance -> the reference's LinAlgError
    m3.init_state_covs[1][0, 0] = np.nan
    with pytest.raises(np.linalg.LinAlgError):
        m3.mle_cluster_assignment()
    # regress / regress_alpha (MMM:1167-1233) against numpy
    rng = np.random.default_rng(9)
    X, B = rng.normal(size=(500, 3)), rng.normal(size=(3, 2))
    Y = X @ B + 0.1 * rng.normal(size=(500, 2))
    for alpha in (0.0, 0.7):
        Bh, Sh = mix.MMLinGaussSS_marginalizable.regress_alpha(X, Y, alpha)
        Bn = np.linalg.solve(X.T @ X + alpha * np.eye(3), X.T @ Y)
        assert relerr(Bh, Bn) < RTOL
        assert relerr(Sh, np.cov(Y - X @ Bn, rowvar=False)) < RTOL


def test_pickle_roundtrip(mix, tmp_path):
    z, x = _synthetic(200, 4, 2, 3, 2, seed=8)
    mdl = mix.MMLinGaussSS_marginalizable(n_clusters=2, states=z, observations=x, random_seed=3).train(n_steps=5)
    mdl.to_pickle(save_location=str(tmp_path))
    mdl.to_pickle(save_location=str(tmp_path))
    files = list(tmp_path.glob(f"mmm-{mdl.hex_hash}*"))
    assert len(files) == 1  # "there can only be one" (MMM:237-245)
    back = mix.MMLinGaussSS_marginalizable.from_pickle(str(files[0]), training_data={"states": z, "observations": x})
    assert back.hex_hash == mdl.hex_hash
    assert np.array_equal(back._packed_params(), mdl._packed_params())
    assert np.array_equal(back.mle_cluster_assignment(), mdl.mle_cluster_assignment())
    # the cache of train_with_multiple_random_starts (MMM:1400-1428, 1467-1468): the second call is a cache hit
    import mtm_b200.framework.marginalizable_mixture_model as mod

    old_home = mod.home_dir
    mod.home_dir = str(tmp_path / "home")
    try:
        proto = mix.MMLinGaussSS_marginalizable(n_clusters=2, states=z, observations=x)
        first = proto.train_with_multiple_random_starts(n_starts=3, n_steps=20, use_cache=True)
        assert len(list((tmp_path / "home" / "tmp").glob(f"mmm-{proto.hex_hash}*"))) == 1
        again = mix.MMLinGaussSS_marginalizable(n_clusters=2, states=z, observations=x).train_with_multiple_random_starts(
            n_starts=3, n_steps=20, use_cache=True)
        assert again is not first and again.last_trained == first.last_trained
        assert np.array_equal(again._packed_params(), first._packed_params())
        assert np.array_equal(again.cluster_assignment, first.cluster_assignment)
    finally:
        mod.home_dir = old_home


def test_kmeans_start_on_device(g_restarts, mix, eng, monkeypatch):
    """init="kmeans" on the device (mtm_kmeans_init): the golden sets' k-means starts (sklearn, MMM:132-151,
    incl. the NaN -> column-mean imputation of set A) are reproduced label for label, and on a large synthetic set
    the device start equals a host sklearn KMeans on the same features."""
    import sklearn.cluster as skl_cluster

    monkeypatch.setenv("MTM_B200_KMEANS", "device")
    for tag in ("A", "B"):
        z, x = g_restarts[f"{tag}_z"], g_restarts[f"{tag}_x"]
        n_c = int(g_restarts[f"{tag}_meta"][4])
        km = mix.MMLinGaussSS_marginalizable(n_clusters=n_c, states=z, observations=x, random_seed=0, init="kmeans")
        assert km.kmeans_info is not None and km.kmeans_info["converged"] in (1, 2)
        assert np.array_equal(km.cluster_assignment, g_restarts[f"{tag}_kmeans_assign0"]), tag
    monkeypatch.setenv("MTM_B200_KMEANS", "auto")
    n, T, d, l, n_c = 300_000, 4, 3, 2, 5
    z, x = _synthetic(n, T, d, l, n_c, seed=13, sep=1.0)
    mdl = mix.MMLinGaussSS_marginalizable(n_clusters=n_c, states=z, observations=x, random_seed=7, init="kmeans")
    assert mdl.kmeans_info is not None
    idx = np.argmax(np.isfinite(z).all(axis=2), axis=0)
    first = z[idx, np.arange(n), :]
    first = np.where(np.isfinite(first), first, np.nanmean(first, axis=0, keepdims=True))
    host = skl_cluster.KMeans(n_clusters=n_c, init="k-means++", random_state=7).fit_predict(first)
    agree = float(np.mean(host == mdl.cluster_assignment))
    assert agree > 0.9999, agree
    assert np.array_equal(np.bincount(host, minlength=n_c) > 0, np.bincount(mdl.cluster_assignment, minlength=n_c) > 0)


def test_fingerprint_cache_key(mix, eng, monkeypatch):
    """mtm_fingerprint: the same key for the same data whatever the storage order, another key after a one-bit
    change; the class uses it as cache key for large data and keeps the reference's md5 on demand."""
    import hashlib

    import mtm_b200

    z, x = _synthetic(70_003, 8, 4, 4, 3, seed=5)
    plain = mtm_b200.DeviceData(z, x, eng.device)
    ordered = mtm_b200.DeviceData(z, x, eng.device, pattern_order=True)
    k0, k1 = eng.fingerprint(plain), eng.fingerprint(ordered)
    assert k0 == k1 and len(k0) == 32
    z2 = z.copy()
    z2[3, 777, 1] = np.nextafter(z2[3, 777, 1], np.inf) if np.isfinite(z2[3, 777, 1]) else 0.0
    assert eng.fingerprint(mtm_b200.DeviceData(z2, x, eng.device)) != k0
    zs = z.copy()
    zs[:, [1, 2]] = zs[:, [2, 1]]  # two trajectories swapped: another data set
    assert eng.fingerprint(mtm_b200.DeviceData(zs, x, eng.device)) != k0
    mdl = mix.MMLinGaussSS_marginalizable(n_clusters=3, states=z, observations=x)
    ref = hashlib.md5(z.tobytes() + x.tobytes() + b"3").hexdigest()
    assert mdl.reference_hex_hash == ref and mdl.hex_hash == ref  # 36 MB: below the threshold, the reference's key
    big_z, big_x = np.tile(z, (1, 2, 1)), np.tile(x, (1, 2, 1))
    big = mix.MMLinGaussSS_marginalizable(n_clusters=3, states=big_z, observations=big_x)
    assert len(big.hex_hash) == 32 and big.hex_hash != big.reference_hex_hash
    again = mix.MMLinGaussSS_marginalizable(n_clusters=3, states=big_z.copy(), observations=big_x.copy())
    assert again.hex_hash == big.hex_hash
    monkeypatch.setenv("MTM_B200_HASH", "md5")
    assert mix.MMLinGaussSS_marginalizable(n_clusters=3, states=big_z, observations=big_x).hex_hash == big.reference_hex_hash


def test_batched_restarts_equal_sequential(mix, eng):
    """mtm_sweep_fit (kernel 4): R restarts fitted in lock-step give, restart by restart, exactly what R
    separate mtm_em_fit calls give (same kernels, one grid row per restart), including restarts that
    stop early while others keep iterating, and the per-restart Q / LL."""
    import mtm_b200

    z, x = _synthetic(3000, 6, 3, 3, 3, seed=21)
    n_c, R, n_steps = 3, 12, 40
    data = mtm_b200.DeviceData(z, x, eng.device)
    a0 = np.stack([np.random.default_rng(100 + r).integers(0, n_c, z.shape[1]) for r in range(R)]).astype(np.int32)
    a0[5, :] = 0
    a0[5, :3] = [1, 2, 2]  # restart 5: near-empty clusters from the start -> guard0, nothing runs
    batch = torch.from_numpy(a0).to(eng.device)
    P, infos, scores, sstat = eng.sweep_fit(data, batch, n_c, 0.0, n_steps)
    exits = set()
    # batched restarts re-sum every cluster in every iteration; a single fit does the same with delta statistics
    # off (bit-identical then) and differs by rounding with them on (the default)
    for delta in (False, True):
        eng.set_delta_stats(2 if delta else 0)
        try:
            for r in range(R):
                single = torch.from_numpy(a0[r].copy()).to(eng.device)
                p1, i1 = eng.em_fit(data, single, n_c, 0.0, n_steps)
                assert i1["exit"] == infos[r]["exit"] and i1["switches"] == infos[r]["switches"], (r, i1, infos[r])
                assert torch.equal(single, batch[r])
                exits.add(i1["exit"])
                if i1["exit"] == "guard0":
                    assert bool(torch.isnan(P[r]).all())
                    continue
                if delta:
                    assert relerr(p1.cpu().numpy(), P[r].cpu().numpy()) < 1e-11
                    continue
                assert torch.equal(p1, P[r])
                q, ll, st = eng.score(data, p1, n_c)
                assert (q, ll) == (scores[r, 0], scores[r, 1]) and st == int(sstat[r])
        finally:
            eng.set_delta_stats(1)
    assert "guard0" in exits and "converged" in exits
    assert len({i["n_e_steps"] for i in infos}) > 2  # restarts really stopped at different iterations


def test_model_selection_sweep(mix):
    """model-selection.py:45-72 on the batched path: AIC/BIC/Q per n_c are those of the selected models."""
    from mtm_b200 import sweep

    z, x = _synthetic(1500, 5, 2, 3, 3, seed=33, sep=0.8)
    table, models = sweep.select_n_clusters(z, x, range(1, 5), n_starts=4, n_steps=30)
    assert list(table["n_clusters"]) == [1, 2, 3, 4]
    for k, n_c in enumerate(range(1, 5)):
        m = models[n_c]
        assert table["n_free_params"][k] == m.n_free_params
        assert abs(table["aic"][k] - (-2 * m.model_log_likelihood() + 2 * m.n_free_params)) < 1e-6
        assert np.isfinite(table["bic"][k]) and np.isfinite(table["Q"][k])
    # more clusters never lower the selected Q on this well-separated 3-cluster data up to the truth
    assert table["Q"][2] > table["Q"][0]


def test_model_selection_matches_reference(mix):
    """model-selection.py:45-72 against the reference's own output (tests/golden/model_selection.npz): for every n_c
    the same winner (seed, labels) and the same AIC / BIC / Q / LL / number of free parameters."""
    from mtm_b200 import sweep

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "model_selection.npz"))
    ref = g["table"]
    table, models = sweep.select_n_clusters(g["z"], g["x"], range(1, 5), n_starts=int(g["n_starts"]), n_steps=100)
    for k, n_c in enumerate(range(1, 5)):
        assert models[n_c].random_seed == int(ref[k, 6]), n_c
        assert np.array_equal(models[n_c].cluster_assignment, g[f"assign_{n_c}"]), n_c
        assert table["n_free_params"][k] == ref[k, 5]
        for col, name in ((1, "aic"), (2, "bic"), (3, "Q"), (4, "LL")):
            assert relerr(table[name][k], ref[k, col]) < RTOL, (n_c, name)


def test_host_streamed_step_equals_resident_step(eng):
    """mtm_em_step_host (slab-pipelined upload) == mtm_em_step on resident data: labels and switch count
    bit-exact, parameters to rounding (the block-partial grouping differs), for ragged slab sizes."""
    import mtm_b200

    n, T, d, l, n_c = 70_003, 6, 4, 4, 5
    z, x = _synthetic(n, T, d, l, n_c, seed=77)
    dev = eng.device
    data = mtm_b200.DeviceData(z, x, dev)
    a0 = np.random.default_rng(3).integers(0, n_c, n).astype(np.int32)
    a_res = torch.from_numpy(a0.copy()).to(dev)
    p_res, _, _ = eng.mstep(data, a_res, n_c, 0.0)
    p_host_dev = p_res.clone()
    sw_res, mc_res, st_res = eng.em_step(data, a_res, p_res, n_c, 0.0)
    zh, xh = torch.from_numpy(z).pin_memory(), torch.from_numpy(x).pin_memory()
    ah = torch.from_numpy(a0.copy()).pin_memory()
    ph = torch.empty(p_res.shape, dtype=torch.float64).pin_memory()
    zd, xd = torch.full_like(data.z, float("nan")), torch.full_like(data.x, float("nan"))
    ad = torch.empty(n, dtype=torch.int32, device=dev)
    for n_slabs in (1, 7, 16):
        ah.copy_(torch.from_numpy(a0))
        pd = p_host_dev.clone()
        sw, mc, st = eng.em_step_host(zh, xh, zd, xd, ah, ad, pd, ph, n_c, 0.0, n_slabs=n_slabs)
        assert (sw, mc, st) == (sw_res, mc_res, st_res)
        assert np.array_equal(ah.numpy(), a_res.cpu().numpy())
        assert relerr(ph.numpy(), p_res.cpu().numpy()) < 1e-12
        assert torch.equal(pd.cpu(), ph)


def test_lg_component_score():
    """framework_extended.StateSpaceLinearGaussian.score on the E-step kernel (single cluster) against
    the reference's score, on the training data and on held-out data with extra missing states."""
    from mtm_b200.framework_extended import state_space_model_linear_gaussian as ssm_lg

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "lg_component.npz"))
    for tag in ("a0", "a1"):
        mdl = ssm_lg.StateSpaceLinearGaussian(alpha=float(g[f"{tag}_alpha"])).fit((g[f"{tag}_z"], g[f"{tag}_x"]))
        assert relerr(mdl.score(), g[f"{tag}_score"]) < RTOL
        assert relerr(mdl.score((g[f"{tag}_zt"], g[f"{tag}_xt"])), g[f"{tag}_score_t"]) < RTOL


def test_lg_component_fit_matches_reference():
    """framework_extended.StateSpaceLinearGaussian.fit on the device (mtm_mstep_ext with one label: nanmean initial
    mean, 1e-6-ridge / Ridge regressions) against the reference's fitted parameters
    (tests/golden/lg_component.npz)."""
    from mtm_b200.framework_extended import state_space_model_linear_gaussian as ssm_lg

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "lg_component.npz"))
    for tag in ("a0", "a1"):
        mdl = ssm_lg.StateSpaceLinearGaussian(alpha=float(g[f"{tag}_alpha"])).fit((g[f"{tag}_z"], g[f"{tag}_x"]))
        for got, key in ((mdl.state_init["mean"], "mean"), (mdl.state_init["cov"], "cov"),
                         (mdl.state_model["coeff"], "A"), (mdl.state_model["covar"], "G"),
                         (mdl.measurement_model["coeff"], "H"), (mdl.measurement_model["covar"], "L")):
            assert relerr(got, g[f"{tag}_{key}"]) < RTOL, (tag, key)
        back = ssm_lg.StateSpaceLinearGaussian().from_pickle(mdl.to_pickle())
        assert np.array_equal(back.state_model["coeff"], mdl.state_model["coeff"]) and back.alpha == mdl.alpha


def test_device_random_labels_bit_exact(eng):
    """mtm_random_labels == np.random.default_rng(seed).integers(0, n_c, n) bit for bit (MMM:166-168),
    including streams in which numpy's Lemire rejection fires (every later element shifts by one draw)."""
    for seed, n_c, n in [(42, 4, 1000), (100, 8, 1_000_003), (7, 3, 65), (0, 1, 10), (123, 12, 5_000_000),
                         (5, 70_000, 20_000_000), (9, 70_000, 129)]:
        got, n_rej = eng.random_labels(seed, n_c, n)
        want = np.random.default_rng(seed).integers(low=0, high=n_c, size=n)
        assert np.array_equal(got.cpu().numpy().astype(np.int64), want), (seed, n_c, n, n_rej)
        if n_c == 70_000 and n > 1_000_000:
            assert n_rej > 50  # the rejection path really ran


@pytest.mark.parametrize("n_c", [12, 24])
def test_many_clusters_vs_oracle(mix, n_c):
    """n_c = 12 (fused kernel at two blocks per SM) and n_c = 24 (statistics too large to fuse: separate
    pass) against the oracle: same switch sequence, labels and parameters."""
    from oracle import cport as oc
    from oracle import dense as od

    n, T, d, l = 40_000, 5, 4, 4
    z, x = _synthetic(n, T, d, l, n_c, seed=n_c, sep=0.6)
    a0 = od.random_init(n, n_c, 3)
    ref = od.train(z, x, a0, n_c, alpha=0.0, n_steps=3, e_step_fn=oc.e_step_seq)
    mdl = mix.MMLinGaussSS_marginalizable(n_clusters=n_c, states=z, observations=x, random_seed=3)
    mdl.train(n_steps=3)
    assert mdl.fit_info["switches"] == ref.switches
    assert np.array_equal(mdl.cluster_assignment, ref.assign)
    assert relerr(mdl._packed_params(), ref.params.packed()) < RTOL


def test_pattern_order(eng):
    """mtm_pattern_order / mtm_permute_*: a stable sort by the per-step "none / one / more hidden-state coordinates
    missing" pattern; gathers exact; inverse round trip; the E-step gives the same bits in either order (point-mass
    and rank-one shortcuts vs generic path)."""
    import mtm_b200

    for (n, T, d, l) in ((70_003, 8, 4, 4), (5_000, 3, 2, 4), (33, 12, 1, 1), (2, 8, 4, 4)):
        z, x = _synthetic(n, T, d, l, 3, seed=11 + n % 7)
        z[:, ::5] = np.where(np.isnan(z[:, ::5]), 0.25, z[:, ::5])  # plenty of complete steps
        z[0, 3 % n, 0] = np.inf
        ordered = mtm_b200.DeviceData(z, x, eng.device, pattern_order=True)
        perm = ordered.perm.cpu().numpy()
        assert np.array_equal(np.sort(perm), np.arange(n))
        key = np.zeros(n, dtype=np.int64)
        for t in range(min(T, 8)):  # base-3 digit per step, step 0 most significant: 2 complete, 1 one missing
            missing = np.sum(~np.isfinite(z[t]), axis=1)
            key = key * 3 + np.where(missing == 0, 2, np.where(missing == 1, 1, 0))
        assert np.array_equal(perm, np.argsort(key, kind="stable"))
        assert np.array_equal(ordered.z.cpu().numpy(), z[:, perm], equal_nan=True)
        assert np.array_equal(ordered.x.cpu().numpy(), x[:, perm], equal_nan=True)
        lab = np.random.default_rng(1).integers(0, 3, (2, n)).astype(np.int32)
        lr = ordered.to_resident(lab)
        assert np.array_equal(lr.cpu().numpy(), lab[:, perm])
        assert np.array_equal(ordered.from_resident(lr).cpu().numpy(), lab)
        if n < 100:
            continue
        plain = mtm_b200.DeviceData(z, x, eng.device)
        mdl_P = eng.mstep(plain, torch.from_numpy(lab[0]).to(eng.device), 3, 0.0)[0]
        r0 = eng.estep(plain, mdl_P, 3, T, want_ll=True, want_probs=True)
        r1 = eng.estep(ordered, mdl_P, 3, T, want_ll=True, want_probs=True)
        assert np.array_equal(ordered.from_resident(r1["ll"]).cpu().numpy(), r0["ll"].cpu().numpy(), equal_nan=True)
        assert np.array_equal(ordered.from_resident(r1["probs"]).cpu().numpy(), r0["probs"].cpu().numpy(),
                              equal_nan=True)
        assert np.array_equal(ordered.from_resident(r1["assign"]).cpu().numpy(), r0["assign"].cpu().numpy())


@pytest.mark.parametrize("shape", [(60_007, 8, 4, 4, 4), (9_001, 5, 2, 3, 3), (5_003, 4, 3, 5, 2)])
def test_delta_statistics(eng, shape):
    """Delta statistics (mtm_b200.h): late EM iterations only visit the trajectories that change cluster.  Stepwise
    EM through mtm_em_step with the mode on and off: the same switch counts and labels at every iteration,
    parameters equal to rounding; the mode is really taken once few labels move; a second step from restored labels
    hits the cache and reproduces the first bit for bit; labels changed behind the library's back miss it."""
    import ctypes

    import mtm_b200
    from mtm_b200 import _capi as C

    n, T, d, l, n_c = shape
    z, x = _synthetic(n, T, d, l, n_c, seed=n % 97, sep=0.6)
    data = mtm_b200.DeviceData(z, x, eng.device, pattern_order=True)
    a0 = data.to_resident(np.random.default_rng(3).integers(0, n_c, n).astype(np.int32))
    vp = lambda t: ctypes.c_void_p(t.data_ptr())

    def step(a, P):
        info = C.EStepInfo()
        C.check(eng.lib.mtm_em_step(eng.h, n_c, 0.0, vp(a), vp(P), 3, ctypes.byref(info)), eng.h)
        return int(info.n_switches), int(info.stats_mode), int(info.status)

    runs = {}
    try:
        for enable in (False, True):  # (changing the mode drops the cache: the delta run goes last and leaves it warm)
            eng.set_delta_stats(2 if enable else 0)  # 2: also for sets this small
            a = a0.clone()
            P = eng.mstep(data, a, n_c, 0.0)[0]  # binds the set once; the steps below do not re-bind
            hist = [step(a, P) for _ in range(14)]
            runs[enable] = (hist, a.clone(), P.clone())
        on, off = runs[True], runs[False]
        assert [h[1] for h in off[0]] == [0] * 14
        assert [h[0] for h in on[0]] == [h[0] for h in off[0]] and all(h[2] == 0 for h in on[0])
        modes = [h[1] for h in on[0]]
        assert modes[0] == 0 and 1 in modes, modes
        for k in range(1, 14):  # delta exactly when the iteration before moved at most n/8 labels
            assert modes[k] == (1 if on[0][k - 1][0] * 8 <= n else 0), (k, modes, on[0])
        assert torch.equal(on[1], off[1])
        assert relerr(on[2].cpu().numpy(), off[2].cpu().numpy()) < 1e-11
        # stationary stepping: restore the labels and parameters, step again -> cache hit, same bits
        a, P = on[1].clone(), on[2].clone()
        a_keep, P_keep = a.clone(), P.clone()
        s1 = step(a, P)
        a1, P1 = a.clone(), P.clone()
        a.copy_(a_keep)
        P.copy_(P_keep)
        s2 = step(a, P)
        assert s1 == s2 and (s2[1] == 1 or s2[0] * 8 > n)
        assert torch.equal(a, a1) and torch.equal(P, P1)
        # labels the caller changed are not in the cache: the iteration sums everything and is right
        a.copy_(a_keep)
        P.copy_(P_keep)
        a[: n // 50] = (a[: n // 50] + 1) % n_c
        b, Q = a.clone(), P.clone()
        s3 = step(a, P)
        assert s3[1] == 0
        eng.set_delta_stats(0)
        s4 = step(b, Q)
        assert s3[0] == s4[0] and torch.equal(a, b) and torch.equal(P, Q)  # both full: identical
    finally:
        eng.set_delta_stats(1)


def test_pattern_order_workspace(eng):
    """The order works in the handle's grow-only workspace: the same permutation from a fresh, a reused (larger set
    first, then a smaller one) and a released workspace."""
    import mtm_b200

    perms = []
    for n in (40_000, 3_001, 40_000):
        z, x = _synthetic(n, 8, 4, 4, 3, seed=5)
        perms.append(mtm_b200.DeviceData(z, x, eng.device, pattern_order=True).perm.cpu().numpy())
        if n < 10_000:
            eng.release_workspace()
    assert np.array_equal(perms[0], perms[2])
    assert np.array_equal(np.sort(perms[1]), np.arange(3_001))
    eng.release_workspace()
    eng.release_workspace()  # idempotent


def test_pattern_order_fit_equals_plain_fit(eng):
    """A whole fit on the pattern-ordered set: same labels and switch history as in the caller's order (the
    statistics are summed in another order, so parameters agree to rounding, not bit for bit)."""
    import mtm_b200

    n, T, d, l, n_c = 200_000, 8, 4, 4, 4
    z, x = _synthetic(n, T, d, l, n_c, seed=21)
    a0 = np.random.default_rng(42).integers(0, n_c, n).astype(np.int32)
    plain = mtm_b200.DeviceData(z, x, eng.device)
    ordered = mtm_b200.DeviceData(z, x, eng.device, pattern_order=True)
    ap = torch.from_numpy(a0).to(eng.device)
    ao = ordered.to_resident(a0)
    pp, ip = eng.em_fit(plain, ap, n_c, 0.0, 6)
    po, io = eng.em_fit(ordered, ao, n_c, 0.0, 6)
    assert ip["switches"] == io["switches"] and ip["exit"] == io["exit"]
    assert np.array_equal(ordered.from_resident(ao).cpu().numpy(), ap.cpu().numpy())
    assert relerr(po.cpu().numpy(), pp.cpu().numpy()) < 1e-11


def test_secondary_predictors(g_pred, mix):
    """SURVEY 8 f3 through the drop-in class: a model trained here from the reference's seed must reproduce
    the reference's fit, and every predictor layered on the E-step must match the reference's outputs."""
    g = g_pred
    ztr, xtr, zte, xte = g["ztr"], g["xtr"], g["zte"], g["xte"]
    n_c, T = g["params"].shape[0], ztr.shape[0]
    mdl = mix.MMLinGaussSS_marginalizable(n_clusters=n_c, states=ztr, observations=xtr, random_seed=int(g["seed"]))
    mdl.train(n_steps=40)
    assert np.array_equal(mdl.cluster_assignment, g["assign"])
    assert relerr(mdl._packed_params(), g["params"]) < RTOL
    assert relerr(mdl.cluster_propensities_over_time(states=ztr, observations=xtr), g["pot_tr"]) < RTOL
    assert relerr(mdl.cluster_propensities_over_time(), g["pot_tr"]) < RTOL  # resident (pattern-ordered) set
    assert relerr(mdl.cluster_propensities_over_time(states=zte, observations=xte), g["pot_te"]) < RTOL
    assert relerr(mdl.observed_cluster_propensities_over_time(), g["opot_tr"]) < RTOL
    assert relerr(mdl.observed_cluster_propensities_over_time(observations=xte), g["opot_te"]) < RTOL
    assert np.array_equal(mdl.initial_full_data_cluster_assignment(), g["init_tr"])
    assert np.array_equal(mdl.initial_full_data_cluster_assignment(states=zte, observations=xte), g["init_te"])
    ps, po = mdl.predictions_from_initial_data()
    assert relerr(ps, g["pfi_s"]) < RTOL and relerr(po, g["pfi_o"]) < RTOL
    s1, o1 = mdl.one_step_ahead_predictions(states=zte, observations=xte)
    assert relerr(s1, g["osa_s"]) < RTOL and relerr(o1, g["osa_o"]) < RTOL
    s2, o2 = mdl.one_step_ahead_predictions_no_history(states=zte, observations=xte)
    assert relerr(s2, g["osanh_s"]) < RTOL and relerr(o2, g["osanh_o"]) < RTOL
    for T0 in (1, 3, T):
        ll = np.stack([mdl.conditional_log_likelihoods_first_T0_steps(c, T0) for c in range(n_c)])
        assert relerr(ll, g[f"ll_T{T0}"]) < RTOL
        assert relerr(mdl.log_likelihood_matrix(T0=T0), g[f"ll_T{T0}"]) < RTOL
        oll = np.stack([mdl.observed_condl_log_lik_first_T0_steps(c, T0, observations=xte) for c in range(n_c)])
        assert relerr(oll, g[f"oll_T{T0}_te"]) < RTOL
    for name, val in (("aic", mdl.aic()), ("bic", mdl.bic()), ("aic_te", mdl.aic(states=zte, observations=xte)),
                      ("bic_te", mdl.bic(states=zte, observations=xte)), ("Q", mdl.e_complete_data_log_lik()),
                      ("LL", mdl.model_log_likelihood())):
        assert abs(val - float(g[name])) < RTOL * abs(float(g[name])), name
    assert mdl.n_free_params == int(g["n_free"])


@pytest.mark.parametrize("shape", [(3001, 8, 4, 4, 3), (700, 5, 2, 3, 4), (257, 4, 6, 2, 2), (129, 12, 1, 1, 2)])
def test_prefix_loglik_one_pass(eng, shape):
    """mtm_estep_prefix: every prefix log-likelihood from one launch equals the per-T0 E-step (same bits), joint
    and observation-only, static and run-time dimensions, pattern-ordered or not."""
    import mtm_b200

    n, T, d, l, n_c = shape
    z, x = _synthetic(n, T, d, l, n_c, seed=31 + n % 5)
    lab = torch.from_numpy(np.random.default_rng(3).integers(0, n_c, n).astype(np.int32)).to(eng.device)
    plain = mtm_b200.DeviceData(z, x, eng.device)
    P = eng.mstep(plain, lab, n_c, 0.0)[0]
    for data, dd in ((plain, None), (mtm_b200.DeviceData(z, x, eng.device, pattern_order=True), None),
                     (mtm_b200.DeviceData(None, x, eng.device), d)):
        pre, status = eng.estep_prefix(data, P, n_c, T, d=dd)
        assert status == 0 and tuple(pre.shape) == (T, n_c, n)
        for T0 in range(1, T + 1):
            ll = eng.estep(data, P, n_c, T0, want_ll=True, d=dd)["ll"]
            assert torch.equal(pre[T0 - 1], ll), (shape, T0)


def test_full_size_properties(eng):
    """BASELINE config 3 at its full size (10 M trajectories, T = 8, d = l = 4, n_c = 8, 20 % missing; generated
    on the device): oracle-free properties of two EM iterations — same labels and switch counts on the
    pattern-ordered set and in the caller's order, cluster sizes add up, the statistics of two shards add up to
    those of the whole, and an oracle check of the log-likelihoods on a strided sample."""
    import bench
    import mtm_b200
    from oracle import cport as oc

    n, T, d, l, n_c = 10_000_000, 8, 4, 4, 8
    z, x = bench.sample_device(n, T, d, l, bench.mixture_params(n_c, d, l), seed=4321, device=eng.device)
    gen = torch.Generator(device=eng.device)
    gen.manual_seed(42)
    a0 = torch.randint(0, n_c, (n,), generator=gen, device=eng.device, dtype=torch.int32)
    plain = mtm_b200.DeviceData(z, x, eng.device)
    ordered = mtm_b200.DeviceData(z, x, eng.device, pattern_order=True)
    ap, ao = a0.clone(), ordered.to_resident(a0)
    pp, ip = eng.em_fit(plain, ap, n_c, 0.0, 2)
    po, io = eng.em_fit(ordered, ao, n_c, 0.0, 2)
    assert ip["switches"] == io["switches"] and ip["n_e_steps"] == 2
    back = ordered.from_resident(ao)
    assert torch.equal(back, ap)
    assert relerr(po.cpu().numpy(), pp.cpu().numpy()) < 1e-11
    sizes = torch.bincount(ap.long(), minlength=n_c)
    assert int(sizes.sum()) == n and io["min_count"] == int(sizes.min())
    # shard additivity of the statistics (what the all-reduce relies on), at full size
    full = eng.mstats(plain, ap, n_c)
    h = n // 2 + 12345
    s1 = eng.mstats(mtm_b200.DeviceData(z[:, :h].contiguous(), x[:, :h].contiguous(), eng.device), ap[:h].contiguous(), n_c)
    s2 = eng.mstats(mtm_b200.DeviceData(z[:, h:].contiguous(), x[:, h:].contiguous(), eng.device), ap[h:].contiguous(), n_c)
    assert relerr((s1 + s2).cpu().numpy(), full.cpu().numpy()) < 1e-12
    assert torch.equal(full[:, 0], sizes.double())
    # oracle on a strided sample of the full-size E-step
    idx = torch.arange(0, n, 4999, device=eng.device)
    sub = mtm_b200.DeviceData(z[:, idx].contiguous(), x[:, idx].contiguous(), eng.device)
    r = eng.estep(sub, pp, n_c, T, want_ll=True)
    P = pp.cpu().numpy()
    llo = oc.seq_loglik_matrix(P, sub.z.cpu().numpy(), sub.x.cpu().numpy())
    assert relerr(r["ll"].cpu().numpy(), llo) < RTOL
    assert np.array_equal(r["assign"].cpu().numpy(), oc.assign(llo, P[:, 0]))


def test_extended_mixture_and_classifier(g_ext, tmp_path):
    """framework_extended's StateSpaceMixtureModel and StateSpaceModelClassifier with linear-Gaussian
    components against the reference's own fits (random start with restarts, k-means-all start) and predictions."""
    import mtm_b200  # noqa: F401
    from mtm_b200.framework_extended import state_space_model_classifier as xc
    from mtm_b200.framework_extended import state_space_model_linear_gaussian as xl
    from mtm_b200.framework_extended import state_space_model_mixture as xm

    g = g_ext
    xm.home_dir = str(tmp_path)  # restarts go through the cache like the reference's (scratch directory)
    ztr, xtr, zte, xte = g["ztr"], g["xtr"], g["zte"], g["xte"]
    for tag, kw in (("rand", dict(init="random", n_iter=50, n_restarts=2)), ("km", dict(init="k-means-all", n_iter=50))):
        zz = ztr if tag == "rand" else np.where(np.isnan(ztr), 0.0, ztr)
        mdl = xm.StateSpaceMixtureModel(n_clusters=2, data=(zz, xtr), component_model=xl.StateSpaceLinearGaussian,
                                        rng=np.random.default_rng(3)).fit(use_cache=False, **kw)
        assert np.array_equal(mdl.cluster_assignment, g[f"{tag}_assign"]), tag
        assert relerr(mdl.cluster_propensities, g[f"{tag}_pi"]) < RTOL
        assert relerr(mdl.cluster_models[0].state_model["coeff"], g[f"{tag}_A0"]) < RTOL
        assert abs(mdl.score() - float(g[f"{tag}_score"])) < RTOL * abs(float(g[f"{tag}_score"]))
        assert abs(mdl.model_log_likelihood() - float(g[f"{tag}_mll"])) < RTOL * abs(float(g[f"{tag}_mll"]))
        assert abs(mdl.score(data=(zte, xte)) - float(g[f"{tag}_score_te"])) < RTOL * abs(float(g[f"{tag}_score_te"]))
        assert relerr(mdl.predict_proba(), g[f"{tag}_proba_tr"]) < RTOL
        pp, pn = mdl.predict_proba(data=(zte, xte), return_prenormalized_log_probs=True)
        assert relerr(pp, g[f"{tag}_proba_te"]) < RTOL and relerr(pn, g[f"{tag}_prenorm_te"]) < RTOL
        assert np.array_equal(mdl.predict(data=(zte, xte), letters=False), g[f"{tag}_pred_te"])
        assert relerr(mdl.cluster_assignment_index(cluster="B", data=(zte, xte)), g[f"{tag}_index_te"]) < RTOL
    # the loop ran on the device with the extended class's own stopping rules
    assert mdl.fit_info["exit"] in ("converged", "guard", "exhausted", "empty") and mdl.fit_info["n_m_steps"] >= 1
    # E/M stepping through the class methods agrees with the device-resident loop
    step = xm.StateSpaceMixtureModel(n_clusters=2, data=(np.where(np.isnan(ztr), 0.0, ztr), xtr),
                                     component_model=xl.StateSpaceLinearGaussian, rng=np.random.default_rng(3))
    step.cluster_assignment = step._initial_labels("k-means-all")
    step._M_step()
    for _ in range(50):
        if step._E_step() == 0:
            break
        step._M_step()
    assert np.array_equal(step.cluster_assignment, g["km_assign"])
    assert relerr(step.cluster_models[0].state_model["coeff"], g["km_A0"]) < RTOL
    clf = xc.StateSpaceModelClassifier(component_model=xl.StateSpaceLinearGaussian).fit(data=(ztr, xtr), labels=g["lab_tr"])
    assert abs(clf.score() - float(g["clf_score"])) < RTOL * abs(float(g["clf_score"]))
    assert abs(clf.score(data=(zte, xte)) - float(g["clf_score_te"])) < RTOL * abs(float(g["clf_score_te"]))
    assert relerr(clf.predict_proba((zte, xte)), g["clf_proba_te"]) < RTOL
    assert np.array_equal(clf.predict((zte, xte)), g["clf_pred_te"])
    assert np.array_equal(clf.predict(), g["clf_pred_tr"])


def test_comm_attach_caller_owned_communicator(eng, g_em):
    """mtm_comm_attach: a caller-owned ncclComm_t (created here straight from libnccl, world size 1) carries the
    per-iteration all-reduce; the fit equals the golden trace, and detaching leaves the communicator alive."""
    import ctypes

    import mtm_b200

    nccl = ctypes.CDLL("libnccl.so.2")
    uid = ctypes.create_string_buffer(128)
    assert nccl.ncclGetUniqueId(uid) == 0

    class Uid(ctypes.Structure):
        _fields_ = [("internal", ctypes.c_char * 128)]

    comm = ctypes.c_void_p()
    nccl.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, Uid, ctypes.c_int]
    assert nccl.ncclCommInitRank(ctypes.byref(comm), 1, Uid.from_buffer_copy(uid.raw), 0) == 0
    try:
        eng.comm_attach(comm.value, 1, 0)
        c = _case(g_em, "c2s")
        data = mtm_b200.DeviceData(c["z"], c["x"], eng.device, pattern_order=True)
        assign = data.to_resident(g_em["c2s_assign0"].astype(np.int32))
        params, info = eng.em_fit(data, assign, c["n_c"], c["alpha"], c["n_steps"])
        assert info["switches"] == list(g_em["c2s_switches"])
        assert np.array_equal(data.from_resident(assign).cpu().numpy(), g_em["c2s_assign"])
        assert relerr(params.cpu().numpy(), g_em["c2s_params"]) < RTOL
    finally:
        eng.comm_destroy()  # detaches only: the communicator is still the caller's
    count = ctypes.c_int(0)
    nccl.ncclCommCount.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)]
    assert nccl.ncclCommCount(comm, ctypes.byref(count)) == 0 and count.value == 1
    nccl.ncclCommDestroy.argtypes = [ctypes.c_void_p]
    assert nccl.ncclCommDestroy(comm) == 0


def test_offset_data_trace(g_cond, mix):
    """Data far from the origin (z + 1e4, x - 3e3; |mean| / spread ~ 1e4): the M-step statistics are taken about
    a pivot near the mean and solved in centred form, so the fit follows the reference's trace — E/M stepping and
    the whole fit — at 1e-9 RELATIVE.  (One-pass raw moments lose ~8 digits here.)"""
    g, key = g_cond, "offset"
    n, T, d, l, n_c, seed, n_steps = (int(v) for v in g[f"{key}_meta"])
    z, x = g[f"{key}_z"], g[f"{key}_x"]
    mdl = mix.MMLinGaussSS_marginalizable(n_clusters=n_c, states=z, observations=x, random_seed=seed)
    assert np.array_equal(mdl.cluster_assignment, g[f"{key}_assign0"])
    hist, ph, sw = g[f"{key}_assign_hist"].astype(np.int64), g[f"{key}_params_hist"], g[f"{key}_switches"]
    mdl.M_step()
    assert relerr(mdl._packed_params(), ph[0]) < RTOL
    for it in range(len(sw)):
        assert mdl.E_step() == int(sw[it])
        assert np.array_equal(mdl.cluster_assignment, hist[it])
        if it + 1 < len(ph):
            mdl.M_step()
            assert relerr(mdl._packed_params(), ph[it + 1]) < RTOL
    fit = mix.MMLinGaussSS_marginalizable(n_clusters=n_c, states=z, observations=x, random_seed=seed).train(n_steps=n_steps)
    assert fit.fit_info["switches"] == list(sw)
    assert np.array_equal(fit.cluster_assignment, g[f"{key}_assign"])
    assert relerr(fit._packed_params(), g[f"{key}_params"]) < RTOL
    assert relerr(fit.log_likelihood_matrix(), g[f"{key}_ll"]) < RTOL
    assert relerr(fit.e_complete_data_log_lik(), g[f"{key}_Q"]) < RTOL
    # batched restarts share the pivot: same fit through mtm_sweep_fit
    import mtm_b200

    eng = mtm_b200.Engine.get()
    data = fit._train_data()
    a = data.to_resident(np.stack([g[f"{key}_assign0"], g[f"{key}_assign0"]]).astype(np.int32))
    P, infos, scores, _ = eng.sweep_fit(data, a, n_c, 0.0, n_steps)
    assert infos[1]["switches"] == list(sw) and relerr(P[1].cpu().numpy(), g[f"{key}_params"]) < RTOL


def test_rank_deficient_regression(g_cond, mix):
    """alpha = 0 with a design that is not of full column rank.  Underdetermined (fewer rows than regressors): the
    reference's lstsq returns the minimum-norm interpolant — matched.  Exactly duplicated columns: the reference
    returns rounding noise of order 1e13 (rcond = machine precision does not cut the ~1e-16 singular value;
    recorded in the golden file); the device returns the minimum-norm solution, checked against numpy's pinv."""
    g = g_cond
    reg = mix.MMLinGaussSS_marginalizable.regress
    B, C = reg(g["reg_under_X"], g["reg_under_Y"])
    assert relerr(B, g["reg_under_B"]) < RTOL
    assert np.abs(C).max() < 1e-12  # exact interpolation: zero up to the rounding of the moment arithmetic
    X, Y = g["reg_dup_X"], g["reg_dup_Y"]
    B, C = reg(X, Y)
    Bmn = np.linalg.pinv(X) @ Y
    assert relerr(B, Bmn) < 1e-9
    assert relerr(C, np.cov(Y - X @ Bmn, rowvar=False)) < 1e-9
    # M_step on data with a duplicated state coordinate: propensities, initial means and covariances are the
    # reference's; the two regressions are minimum norm
    z, x = g["dupcol_z"], g["dupcol_x"]
    mdl = mix.MMLinGaussSS_marginalizable(n_clusters=2, states=z, observations=x, random_seed=5)
    assert np.array_equal(mdl.cluster_assignment, g["dupcol_assign0"])
    mdl.M_step()
    d = z.shape[2]
    P, R = mdl._packed_params(), g["dupcol_params"]
    assert relerr(P[:, :1 + d + d * d], R[:, :1 + d + d * d]) < RTOL
    for c in range(2):
        Z = z[:, mdl.cluster_assignment == c]
        prev, nxt = Z[:-1].reshape(-1, d), Z[1:].reshape(-1, d)
        ok = np.isfinite(prev).all(axis=1) & np.isfinite(nxt).all(axis=1)
        assert relerr(mdl.transition_matrices[c], np.linalg.pinv(prev[ok]) @ nxt[ok]) < 1e-9
