"""Parity checks of the drop-in class when it shards itself over a torch.distributed group: every rank builds the
model from the same full arrays (like the reference's single-process script, inference-adni-xval.py:127-133) and
must reproduce the reference's golden traces.  Used by tests/test_gpu_dist.py (two spawned ranks) and by
`bench.py --gpus N` (N > 1) as its class-level parity gate."""

import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL, FLOOR = 1e-9, 1e-6


def relerr(a, b):
    a, b = np.atleast_1d(np.asarray(a, dtype=float)), np.atleast_1d(np.asarray(b, dtype=float))
    assert a.shape == b.shape, (a.shape, b.shape)
    with np.errstate(all="ignore"):
        d = np.abs(a - b) / np.maximum(FLOOR, np.abs(b))
    d[np.isnan(a) & np.isnan(b)] = 0.0
    d[a == b] = 0.0
    return float(np.max(d)) if d.size else 0.0


def class_checks(mix, world, rank):
    """Run on every rank of an initialised process group; returns a list of failure strings (empty = pass)."""
    from mtm_b200 import dist as mdist
    from mtm_b200.engine import Engine

    bad = []

    def check(cond, what):
        if not cond:
            bad.append(what)

    g = np.load(os.path.join(ROOT, "tests", "golden", "em_traces.npz"))
    for key in ("c2s", "adnish", "guardexit"):
        n, T, d, l, n_c, seed, n_steps = (int(v) for v in g[f"{key}_meta"])
        z, x, alpha = g[f"{key}_z"], g[f"{key}_x"], float(g[f"{key}_alpha"])
        mdl = mix.MMLinGaussSS_marginalizable(n_clusters=n_c, states=z, observations=x, random_seed=seed, alpha=alpha)
        if n >= 32 * world:
            check(mdl._world == world and mdl._train_data().shard is not None, f"{key}: model is not sharded")
            check(mdl._train_data().n < n, f"{key}: rank holds the whole set")
        else:  # too small to shard: every rank fits its own full copy, without any collective
            check(mdl._world == 1 and mdl._train_data().shard is None, f"{key}: small model was sharded")
        mdl.train(n_steps=n_steps)
        check(mdl.fit_info["switches"] == list(g[f"{key}_switches"]), f"{key}: switch history")
        check(np.array_equal(mdl.cluster_assignment, g[f"{key}_assign"]), f"{key}: labels")
        check(relerr(mdl._packed_params(), g[f"{key}_params"]) < RTOL, f"{key}: parameters")
        if key == "guardexit":
            continue
        a, probs, prenorm = mdl.mle_cluster_assignment(return_probs=True, return_prenormalized_log_probs=True)
        check(np.array_equal(a, g[f"{key}_mle_assign"]), f"{key}: mle labels")
        check(relerr(probs, g[f"{key}_probs"]) < RTOL and relerr(prenorm, g[f"{key}_prenorm"]) < RTOL, f"{key}: probs")
        check(relerr(mdl.e_complete_data_log_lik(), g[f"{key}_Q"]) < RTOL, f"{key}: Q")
        check(relerr(mdl.model_log_likelihood(), g[f"{key}_LL"]) < RTOL, f"{key}: LL")
        check(relerr(mdl.bic(), g[f"{key}_bic"]) < RTOL, f"{key}: BIC")
        zt, xt = z[: max(1, T - 1), : n // 3], x[: max(1, T - 1), : n // 3]
        a, pr, pn = mdl.mle_cluster_assignment(states=zt, observations=xt, return_probs=True,
                                               return_prenormalized_log_probs=True)
        check(np.array_equal(a, g[f"{key}_ho_assign"]) and relerr(pr, g[f"{key}_ho_probs"]) < RTOL, f"{key}: held-out")
        ao, po = mdl.observations_mle_cluster_assignment(return_probs=True)
        check(np.array_equal(ao, g[f"{key}_obs_assign"]) and relerr(po, g[f"{key}_obs_probs"]) < RTOL, f"{key}: obs-only")
        # E_step() / M_step() one at a time
        one = mix.MMLinGaussSS_marginalizable(n_clusters=n_c, states=z, observations=x, random_seed=seed, alpha=alpha)
        one.M_step()
        check(relerr(one._packed_params(), g[f"{key}_params_hist"][0]) < RTOL, f"{key}: first M_step")
        check(one.E_step() == int(g[f"{key}_switches"][0]), f"{key}: first E_step switch count")
        check(np.array_equal(one.cluster_assignment, g[f"{key}_assign_hist"][0].astype(np.int64)), f"{key}: first E_step")
    # restart selection (batched restarts, all-reduced scores)
    gr = np.load(os.path.join(ROOT, "tests", "golden", "restarts.npz"))
    z, x = gr["B_z"], gr["B_x"]
    n, T, d, l, n_c, n_starts = (int(v) for v in gr["B_meta"])
    proto = mix.MMLinGaussSS_marginalizable(n_clusters=n_c, states=z, observations=x)
    best, obj = proto.train_with_multiple_random_starts(n_starts=n_starts, n_steps=100, use_cache=False,
                                                        return_objectives=True)
    check(relerr(obj, gr["B_objectives"]) < RTOL, "restarts: objectives")
    check(best.random_seed == int(gr["B_best_seed"]) and np.array_equal(best.cluster_assignment, gr["B_best_assign"]),
          "restarts: winner")
    check(relerr(best._packed_params(), gr["B_best_params"]) < RTOL, "restarts: winner parameters")
    # detach and re-attach the communicator (per fold in a cross-validation): the peer-memory mailbox and the
    # sequence numbers left in its flags survive, the next exchanges must not match them
    eng = Engine.get_sharded(None, None)
    mdist.detach(eng)
    mdist.attach(eng)
    key = "c2s"
    n, T, d, l, n_c, seed, n_steps = (int(v) for v in g[f"{key}_meta"])
    again = mix.MMLinGaussSS_marginalizable(n_clusters=n_c, states=g[f"{key}_z"], observations=g[f"{key}_x"],
                                            random_seed=seed, alpha=float(g[f"{key}_alpha"])).train(n_steps=n_steps)
    check(again.fit_info["switches"] == list(g[f"{key}_switches"]), "re-attach: switch history")
    check(relerr(again._packed_params(), g[f"{key}_params"]) < RTOL, "re-attach: parameters")
    return bad


