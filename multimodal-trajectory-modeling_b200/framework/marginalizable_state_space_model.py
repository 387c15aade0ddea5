"""Parameter-only helpers with the names of the reference's
``framework/marginalizable_state_space_model.py`` (SSM): the stacked mean and the dense joint
covariance of (Z_1..Z_T, X_1..X_T).

In the reference these feed the dense marginalising log-density (SSM:728-773) and are rebuilt
for every cluster in every E-step.  The CUDA path never forms them (it runs the equivalent
prediction-error recursion in registers, see csrc/filter.cuh); they are kept — as closed-form
numpy on (d, l, T)-sized inputs, nothing per trajectory — for callers that inspect the model's
joint Gaussian.  `multivariate_normal_log_likelihood` for an ARBITRARY covariance is not part of
the mixture path and is not provided; the state-space-structured case is
`MMLinGaussSS_marginalizable.log_likelihood_matrix`.
"""

from __future__ import annotations

import numpy as np


def _marginal_state_covs(T, S, A, G):
    V = [S]
    for _ in range(1, T):
        V.append(G + A.T @ V[-1] @ A)  # SSM:62-64
    return V


def mmZ(T, m, A):
    """E[Z_1..Z_T] = [m, mA, ..] (SSM:294-314)."""
    A = np.atleast_2d(A)
    out, cur = [], np.atleast_1d(np.asarray(m, dtype=float))
    for _ in range(T):
        out.append(cur)
        cur = cur @ A
    return np.concatenate(out)


def mmX(T, m, A, H):
    """E[X_1..X_T] (SSM:317-339)."""
    H = np.atleast_2d(H)
    d = np.atleast_1d(m).shape[0]
    return (mmZ(T, m, A).reshape(T, d) @ H).ravel()


def mm(T, m, A, H):
    """E[Z, X] (SSM:342-364)."""
    return np.concatenate([mmZ(T, m, A), mmX(T, m, A, H)])


def CZZ(T, S, A, G):
    """Var(Z): block (i, j) = V_i A^(j-i), j >= i (SSM:67-133)."""
    S, A, G = map(np.atleast_2d, (S, A, G))
    d = A.shape[0]
    V = _marginal_state_covs(T, S, A, G)
    out = np.empty((T * d, T * d))
    for i in range(T):
        blk = V[i]
        for j in range(i, T):
            out[i * d:(i + 1) * d, j * d:(j + 1) * d] = blk
            out[j * d:(j + 1) * d, i * d:(i + 1) * d] = blk.T
            blk = blk @ A
    return out


def CXX(T, S, A, G, H, L):
    """Var(X) = (I (x) H)' Var(Z) (I (x) H) + I (x) Lambda (SSM:169-256)."""
    S, A, G, H, L = map(np.atleast_2d, (S, A, G, H, L))
    K = np.kron(np.eye(T), H)
    return K.T @ CZZ(T, S, A, G) @ K + np.kron(np.eye(T), L)


def CC(T, S, A, G, H, L):
    """Var([Z, X]) (SSM:259-291)."""
    S, A, G, H, L = map(np.atleast_2d, (S, A, G, H, L))
    zz = CZZ(T, S, A, G)
    K = np.kron(np.eye(T), H)
    zx = zz @ K
    return np.block([[zz, zx], [zx.T, K.T @ zx + np.kron(np.eye(T), L)]])
