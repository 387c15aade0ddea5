"""CPU oracle (numpy): dense restatement of the reference's hard-assignment EM hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this module;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may.

Citation tags (paths relative to the reference checkout):
  SSM = framework/marginalizable_state_space_model.py
  MMM = framework/marginalizable_mixture_model.py

The arithmetic follows the *reference's own formulation* (dense (d+l)T joint
Gaussian, per-row marginalisation by sub-setting the covariance, determinant and
linear solve by LU) so that it is an independent check of the sequential
(prediction-error) formulation used by the CUDA kernels.

Parity status: PINNED.  ``tests/golden/make_golden.py`` runs the unmodified
reference (imported from /root/reference in the build container) and stores its
outputs under ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks
every function here against those vectors.
"""

from __future__ import annotations

import numpy as np

EPS = np.finfo(float).eps

# --------------------------------------------------------------------------
# joint mean / covariance of (Z_1..Z_T, X_1..X_T)            SSM:32-364
# --------------------------------------------------------------------------


def state_marginal_covs(T, S, A, G):
    """V_1 = S, V_t = G + A' V_{t-1} A  (SSM:62-64)."""
    S, A, G = (np.atleast_2d(np.asarray(v, dtype=float)) for v in (S, A, G))
    V = [S]
    for _ in range(1, T):
        V.append(G + A.T @ V[-1] @ A)
    return V


def joint_mean(T, m, A, H):
    """[m, mA, .., mA^{T-1} | mH, mAH, ..]  (SSM:312-314, 337-339, 364)."""
    A, H = (np.atleast_2d(np.asarray(v, dtype=float)) for v in (A, H))
    m = np.atleast_1d(np.asarray(m, dtype=float))
    zs = [m @ np.linalg.matrix_power(A, t) for t in range(T)]
    return np.concatenate(zs + [z @ H for z in zs])


def obs_mean(T, m, A, H):
    """Mean of (X_1..X_T) only (SSM:317-339)."""
    d = np.atleast_1d(m).shape[0]
    return joint_mean(T, m, A, H)[T * d:]


def state_cov(T, S, A, G):
    """Var(Z_1..Z_T): block (i,j) = V_i A^{j-i} for j>i, transpose below (SSM:100-133)."""
    S, A, G = (np.atleast_2d(np.asarray(v, dtype=float)) for v in (S, A, G))
    d = A.shape[0]
    V = state_marginal_covs(T, S, A, G)
    out = np.empty((T * d, T * d))
    for i in range(T):
        for j in range(T):
            if i == j:
                blk = V[i]
            elif j > i:
                blk = V[i] @ np.linalg.matrix_power(A, j - i)
            else:
                blk = (V[j] @ np.linalg.matrix_power(A, i - j)).T
            out[i * d:(i + 1) * d, j * d:(j + 1) * d] = blk
    return out


def obs_cov(T, S, A, G, H, Lam):
    """Var(X_1..X_T): H' Cov(Z_i,Z_j) H (+Lam on the diagonal)  (SSM:217-256).

    Block (i,j) for j<i is the transpose of block (j,i), as in the reference.
    """
    S, A, G, H, Lam = (np.atleast_2d(np.asarray(v, dtype=float)) for v in (S, A, G, H, Lam))
    d, l = H.shape
    V = state_marginal_covs(T, S, A, G)
    out = np.empty((T * l, T * l))
    for i in range(T):
        for j in range(i, T):
            if i == j:
                blk = Lam + H.T @ V[i] @ H
            else:
                blk = H.T @ (V[i] @ np.linalg.matrix_power(A, j - i)) @ H
            out[i * l:(i + 1) * l, j * l:(j + 1) * l] = blk
            if j > i:
                out[j * l:(j + 1) * l, i * l:(i + 1) * l] = blk.T
    return out


def joint_cov(T, S, A, G, H, Lam):
    """[[CZZ, CZX],[CZX', CXX]]  (SSM:259-291); CZX block (i,j) = Cov(Z_i,Z_j) H (SSM:161-166)."""
    S, A, G, H, Lam = (np.atleast_2d(np.asarray(v, dtype=float)) for v in (S, A, G, H, Lam))
    d, l = H.shape
    zz = state_cov(T, S, A, G)
    zx = np.empty((T * d, T * l))
    for i in range(T):
        for j in range(T):
            zx[i * d:(i + 1) * d, j * l:(j + 1) * l] = zz[i * d:(i + 1) * d, j * d:(j + 1) * d] @ H
    return np.block([[zz, zx], [zx.T, obs_cov(T, S, A, G, H, Lam)]])


# --------------------------------------------------------------------------
# NaN/inf-marginalising Gaussian log-density                 SSM:728-773
# --------------------------------------------------------------------------


def marginal_logpdf_rows(Y, mean, cov):
    """log N(y_obs; mean_obs, cov_obs,obs) for each row; non-finite entries of
    (y - mean) are marginalised out (SSM:768-773).  (2*pi)^k * det is formed
    before the log, exactly as the reference does."""
    Y = np.atleast_2d(np.asarray(Y, dtype=float))
    mean = np.asarray(mean, dtype=float).ravel()
    cov = np.atleast_2d(np.asarray(cov, dtype=float))
    out = np.empty(Y.shape[0])
    with np.errstate(all="ignore"):
        for i in range(Y.shape[0]):
            r = Y[i] - mean
            idx = np.flatnonzero(np.isfinite(r))
            sub = cov[np.ix_(idx, idx)]
            ro = r[idx]
            if idx.size and not np.isfinite(sub).all():
                # numba's det/solve refuse non-finite input (numba/np/linalg.py
                # _check_finite_matrix); the reference therefore raises here
                raise np.linalg.LinAlgError("Array must not contain infs or NaNs.")
            det = np.linalg.det(sub) if idx.size else 1.0
            sol = np.linalg.solve(sub, ro) if idx.size else ro
            out[i] = -0.5 * np.log((2.0 * np.pi) ** idx.size * det) - 0.5 * (ro @ sol)
    return out


def stack_rows(states, observations, T0=None):
    """(n, (d+l)T0) matrix [z_1|..|z_T0|x_1|..|x_T0] per trajectory (MMM:490)."""
    T0 = states.shape[0] if T0 is None else T0
    return np.hstack([*states[:T0], *observations[:T0]])


# --------------------------------------------------------------------------
# parameter container
# --------------------------------------------------------------------------


class Params:
    """Mixture parameters; lists indexed by cluster, like the reference's attributes
    (MMM:83-125)."""

    def __init__(self, pi, m, S, A, G, H, Lam):
        self.pi = np.asarray(pi, dtype=float)
        self.m = [np.atleast_1d(np.asarray(v, dtype=float)) for v in m]
        self.S = [np.atleast_2d(np.asarray(v, dtype=float)) for v in S]
        self.A = [np.atleast_2d(np.asarray(v, dtype=float)) for v in A]
        self.G = [np.atleast_2d(np.asarray(v, dtype=float)) for v in G]
        self.H = [np.atleast_2d(np.asarray(v, dtype=float)) for v in H]
        self.Lam = [np.atleast_2d(np.asarray(v, dtype=float)) for v in Lam]

    @property
    def n_clusters(self):
        return len(self.m)

    def copy(self):
        return Params(self.pi.copy(), *[[a.copy() for a in lst] for lst in
                                        (self.m, self.S, self.A, self.G, self.H, self.Lam)])

    def packed(self):
        """(n_c, 1+d+3d^2+d*l+l^2) row-major [pi, m, S, A, G, H, Lam] — the C-ABI layout."""
        rows = []
        for c in range(self.n_clusters):
            rows.append(np.concatenate([[self.pi[c]], self.m[c].ravel(), self.S[c].ravel(),
                                        self.A[c].ravel(), self.G[c].ravel(),
                                        self.H[c].ravel(), self.Lam[c].ravel()]))
        return np.ascontiguousarray(np.stack(rows))

    @staticmethod
    def from_packed(P, d, l):
        P = np.atleast_2d(np.asarray(P, dtype=float))
        pi, m, S, A, G, H, Lam = [], [], [], [], [], [], []
        for row in P:
            o = 0
            pi.append(row[o]); o += 1
            m.append(row[o:o + d].copy()); o += d
            S.append(row[o:o + d * d].reshape(d, d).copy()); o += d * d
            A.append(row[o:o + d * d].reshape(d, d).copy()); o += d * d
            G.append(row[o:o + d * d].reshape(d, d).copy()); o += d * d
            H.append(row[o:o + d * l].reshape(d, l).copy()); o += d * l
            Lam.append(row[o:o + l * l].reshape(l, l).copy()); o += l * l
        return Params(pi, m, S, A, G, H, Lam)

    def n_free_params(self):
        """MMM:197-214."""
        tri = lambda M: M.shape[0] * (M.shape[0] + 1) // 2
        return (self.pi.size + sum(v.size for v in self.m) + sum(v.size for v in self.A)
                + sum(v.size for v in self.H)
                + sum(tri(v) for v in self.S) + sum(tri(v) for v in self.G)
                + sum(tri(v) for v in self.Lam))


# --------------------------------------------------------------------------
# E-step pieces                                              MMM:439-530, 724-791, 997-1165
# --------------------------------------------------------------------------


def cond_loglik(par: Params, c, states, observations, T0=None):
    """ll[i] = log p(z_i, x_i | cluster c) on the first T0 steps (MMM:439-494)."""
    T0 = states.shape[0] if T0 is None else min(T0, states.shape[0])
    mu = joint_mean(T0, par.m[c], par.A[c], par.H[c])
    Sg = joint_cov(T0, par.S[c], par.A[c], par.G[c], par.H[c], par.Lam[c])
    return marginal_logpdf_rows(stack_rows(states, observations, T0), mu, Sg)


def obs_cond_loglik(par: Params, c, observations, T0=None):
    """ll[i] = log p(x_i | cluster c), hidden states integrated out (MMM:997-1048)."""
    T0 = observations.shape[0] if T0 is None else min(T0, observations.shape[0])
    mu = obs_mean(T0, par.m[c], par.A[c], par.H[c])
    Sg = obs_cov(T0, par.S[c], par.A[c], par.G[c], par.H[c], par.Lam[c])
    return marginal_logpdf_rows(np.hstack([*observations[:T0]]), mu, Sg)


def loglik_matrix(par: Params, states, observations, T0=None):
    """(n_c, n) matrix of conditional log-likelihoods."""
    return np.stack([cond_loglik(par, c, states, observations, T0) for c in range(par.n_clusters)])


def obs_loglik_matrix(par: Params, observations, T0=None):
    return np.stack([obs_cond_loglik(par, c, observations, T0) for c in range(par.n_clusters)])


def assign_from_loglik(pi, ll, want_probs=False, want_prenorm=False):
    """Linear-domain argmax of pi_c * exp(ll_c) (first max wins, NaN wins: np.argmax),
    optional probs = L / sum L and prenorm = log pi + ll  (MMM:760-791)."""
    with np.errstate(all="ignore"):
        L = np.asarray(pi)[:, None] * np.exp(ll)
        a = np.argmax(L, axis=0)
        out = [a]
        if want_probs or want_prenorm:
            out.append(L / np.sum(L, axis=0, keepdims=True))
        if want_prenorm:
            out.append(np.log(np.asarray(pi))[:, None] + ll)
    return out[0] if len(out) == 1 else tuple(out)


def e_step(par: Params, states, observations, old_assign):
    """new assignment and number of switches (MMM:1235-1252)."""
    new = assign_from_loglik(par.pi, loglik_matrix(par, states, observations))
    return new, int(np.sum(new != old_assign))


def q_value(par: Params, states, observations, ll=None):
    """expected complete-data log-likelihood Q (MMM:579-624)."""
    ll = loglik_matrix(par, states, observations) if ll is None else ll
    a = assign_from_loglik(par.pi, ll)
    with np.errstate(all="ignore"):
        return float(np.sum(np.log(par.pi[a])) + np.sum(ll[a, np.arange(a.size)]))


def model_loglik(par: Params, states, observations, ll=None):
    """sum_i log sum_c pi_c exp(ll_ci), no log-sum-exp (MMM:626-666)."""
    ll = loglik_matrix(par, states, observations) if ll is None else ll
    with np.errstate(all="ignore"):
        return float(np.sum(np.log(np.sum(par.pi[:, None] * np.exp(ll), axis=0))))


def aic(par, states, observations, ll=None):
    """MMM:668-693."""
    return -2.0 * model_loglik(par, states, observations, ll) + 2 * par.n_free_params()


def bic(par, states, observations, ll=None):
    """MMM:695-722."""
    return (-2.0 * model_loglik(par, states, observations, ll)
            + np.log(states.shape[1]) * par.n_free_params())


# --------------------------------------------------------------------------
# secondary predictors built on the E-step                   MMM:532-577, 808-995, 1077-1123
# --------------------------------------------------------------------------


def propensities_over_time(par: Params, states, observations):
    """(T, n, n_c): membership probabilities from the first t+1 steps (MMM:532-577)."""
    T = states.shape[0]
    with np.errstate(all="ignore"):
        pc = np.stack([(par.pi[:, None] * np.exp(loglik_matrix(par, states, observations, t + 1))).T for t in range(T)])
        return pc / np.sum(pc, axis=-1, keepdims=True)


def obs_propensities_over_time(par: Params, observations):
    """(T, n, n_c): the same from the measurements alone (MMM:1077-1123)."""
    with np.errstate(all="ignore"):
        pc = np.stack([(par.pi[:, None] * np.exp(obs_loglik_matrix(par, observations, t + 1))).T
                       for t in range(observations.shape[0])])
        return pc / np.sum(pc, axis=-1, keepdims=True)


def initial_assignment(par: Params, states, observations):
    """argmax from the first time step only (MMM:909-941)."""
    return assign_from_loglik(par.pi, loglik_matrix(par, states, observations, 1))


def predictions_from_initial(par: Params, states, observations):
    """Noise-free roll-out z_1 A^t, z_1 A^t H under the cluster chosen from the first step (MMM:943-995)."""
    a = initial_assignment(par, states, observations)
    ps, po = np.zeros_like(states), np.zeros_like(observations)
    for i in range(states.shape[1]):
        zt = states[0, i, :]
        for t in range(max(ps.shape[0], po.shape[0])):
            if t < ps.shape[0]:
                ps[t, i, :] = zt
            if t < po.shape[0]:
                po[t, i, :] = zt @ par.H[a[i]]
            zt = zt @ par.A[a[i]]
    return ps, po


def one_step_ahead(par: Params, states, observations, history=True):
    """Probability-weighted one-step-ahead means from the last row (MMM:808-907); history=False keeps only
    the last step when computing the membership probabilities."""
    if history:
        s0, o0 = states, observations
    else:
        s0, o0 = np.nan * np.ones_like(states), np.nan * np.ones_like(observations)
        s0[-1], o0[-1] = states[-1], observations[-1]
    probs = assign_from_loglik(par.pi, loglik_matrix(par, s0, o0), want_probs=True)[1]
    nxt_s = np.zeros((1, *states.shape[1:]))
    nxt_o = np.zeros((1, *observations.shape[1:]))
    for c in range(par.n_clusters):
        s_c = states[-1] @ par.A[c]
        nxt_s += probs[c][:, None] * s_c
        nxt_o += probs[c][:, None] * (s_c @ par.H[c])
    return nxt_s, nxt_o


# --------------------------------------------------------------------------
# M-step                                                     MMM:1167-1308
# --------------------------------------------------------------------------


def _sample_cov(R):
    """np.cov(R, rowvar=False): mean-subtracted, divided by N-1 (MMM:1197-1199, 1230-1232, 1270)."""
    with np.errstate(all="ignore"):
        return np.atleast_2d(np.cov(R, rowvar=False))


def fit_linear(X, Y, alpha):
    """Coefficients of Y ~ X B with no intercept and residual covariance.

    alpha <= 2 eps: least squares (np.linalg.lstsq, MMM:1194); otherwise ridge via the
    normal equations (X'X + alpha I) B = X'Y, which is what sklearn's
    Ridge(solver='auto', fit_intercept=False) does for dense input (MMM:1225-1229)."""
    if X.shape[0] == 0:
        # numba's lstsq raises 'Arrays cannot be empty'; sklearn's Ridge raises ValueError
        raise np.linalg.LinAlgError("Arrays cannot be empty")
    if alpha > 2 * EPS:
        p = X.shape[1]
        B = np.linalg.solve(X.T @ X + alpha * np.eye(p), X.T @ Y)
    else:
        B = np.linalg.lstsq(X, Y, rcond=-1)[0]
    return B, _sample_cov(Y - X @ B)


def m_step(states, observations, assign, n_clusters, alpha=0.0):
    """Parameters from a hard assignment (MMM:1254-1308)."""
    T, n, d = states.shape
    l = observations.shape[2]
    pi = np.empty(n_clusters)
    m, S, A, G, H, Lam = [], [], [], [], [], []
    with np.errstate(all="ignore"):
        for c in range(n_clusters):
            sel = assign == c
            pi[c] = np.mean(sel)                                   # MMM:1261-1263
            Z = states[:, sel, :]
            X = observations[:, sel, :]
            z0 = Z[0][np.isfinite(Z[0]).all(axis=1)]               # MMM:1267-1268
            m.append(np.mean(z0, axis=0))
            S.append(_sample_cov(z0))
            prev = Z[:-1].reshape(-1, d)                           # time-major stacking, MMM:1272-1273
            nxt = Z[1:].reshape(-1, d)
            ok = np.isfinite(prev).all(axis=1) & np.isfinite(nxt).all(axis=1)
            a_c, g_c = fit_linear(prev[ok], nxt[ok], alpha)
            A.append(a_c); G.append(g_c)
            zs = Z.reshape(-1, d)
            xs = X.reshape(-1, l)
            ok = np.isfinite(zs).all(axis=1) & np.isfinite(xs).all(axis=1)   # MMM:1292-1296
            h_c, l_c = fit_linear(zs[ok], xs[ok], alpha)
            H.append(h_c); Lam.append(l_c)
    return Params(pi, m, S, A, G, H, Lam)


# --------------------------------------------------------------------------
# EM loop and multi-restart                                  MMM:1310-1471
# --------------------------------------------------------------------------


def min_cluster_size(assign, n_clusters):
    return int(np.min(np.bincount(assign, minlength=n_clusters)))


class FitResult:
    def __init__(self):
        self.params = None          # None if the initial guard fired (no M-step ever ran)
        self.assign = None
        self.n_e_steps = 0          # number of E-steps executed
        self.switches = []          # n_switches after each E-step
        self.exit = ""              # "guard0" | "converged" | "guard" | "exhausted"


def train(states, observations, assign0, n_clusters, alpha=0.0, n_steps=1000,
          e_step_fn=None, m_step_fn=None):
    """guard -> M -> { E ; stop if 0 switches ; guard ; M }  (MMM:1310-1366).

    ``e_step_fn(par, states, observations, old) -> (new, n_switches)`` and
    ``m_step_fn(states, observations, assign, n_c, alpha) -> Params`` may be replaced by
    faster restatements (the C oracle) in tests."""
    e_step_fn = e_step_fn or e_step
    m_step_fn = m_step_fn or m_step
    alpha = alpha if alpha > 2 * EPS else 0.0                      # MMM:130
    res = FitResult()
    res.assign = np.asarray(assign0).copy()
    if min_cluster_size(res.assign, n_clusters) <= 3:              # MMM:1328-1336
        res.exit = "guard0"
        return res
    res.params = m_step_fn(states, observations, res.assign, n_clusters, alpha)
    res.exit = "exhausted"
    for _ in range(n_steps):
        res.assign, sw = e_step_fn(res.params, states, observations, res.assign)
        res.n_e_steps += 1
        res.switches.append(sw)
        if sw == 0:                                                # MMM:1342-1345
            res.exit = "converged"
            break
        if min_cluster_size(res.assign, n_clusters) <= 3:          # MMM:1346-1356
            res.exit = "guard"
            break
        res.params = m_step_fn(states, observations, res.assign, n_clusters, alpha)
    return res


def random_init(n, n_clusters, seed):
    """rng.integers(0, n_c, n) with default_rng(seed) (MMM:128, 166-168)."""
    return np.random.default_rng(seed).integers(low=0, high=n_clusters, size=n)
