"""ctypes loader for the C oracle (oracle/mtm_oracle.c).  TEST INFRASTRUCTURE ONLY."""

from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from . import dense

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libmtm_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(HERE, "mtm_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE])
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB)
        dp = ctypes.POINTER(ctypes.c_double)
        _lib.orc_dense_loglik.argtypes = [dp, ctypes.c_long, ctypes.c_int, dp, dp, dp, ctypes.c_int]
        _lib.orc_dense_loglik.restype = ctypes.c_int
        _lib.orc_seq_loglik.argtypes = [dp, dp, ctypes.c_int, ctypes.c_long, ctypes.c_int, ctypes.c_int,
                                        dp, ctypes.c_int, ctypes.c_int, dp, ctypes.c_int]
        _lib.orc_seq_loglik.restype = ctypes.c_int
        _lib.orc_assign.argtypes = [dp, dp, ctypes.c_int, ctypes.c_long, ctypes.POINTER(ctypes.c_int64)]
        _lib.orc_assign.restype = None
        _lib.orc_num_threads.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def num_threads():
    return int(lib().orc_num_threads())


def dense_logpdf_rows(Y, mean, cov, nthreads=0):
    """C twin of dense.marginal_logpdf_rows (reference formulation, SSM:728-773)."""
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    mean = np.ascontiguousarray(mean, dtype=np.float64)
    cov = np.ascontiguousarray(cov, dtype=np.float64)
    out = np.empty(Y.shape[0])
    st = lib().orc_dense_loglik(_p(Y), Y.shape[0], Y.shape[1], _p(mean), _p(cov), _p(out), nthreads)
    if st == 1:
        raise np.linalg.LinAlgError("Array must not contain infs or NaNs.")
    if st == 2:
        raise np.linalg.LinAlgError("Matrix is singular to machine precision.")
    return out


def dense_loglik_matrix(par: dense.Params, states, observations, T0=None, nthreads=0):
    """(n_c, n) log-likelihoods, dense formulation, C inner loop."""
    T0 = states.shape[0] if T0 is None else min(T0, states.shape[0])
    Y = dense.stack_rows(states, observations, T0)
    return np.stack([
        dense_logpdf_rows(Y, dense.joint_mean(T0, par.m[c], par.A[c], par.H[c]),
                          dense.joint_cov(T0, par.S[c], par.A[c], par.G[c], par.H[c], par.Lam[c]), nthreads)
        for c in range(par.n_clusters)])


def dense_obs_loglik_matrix(par: dense.Params, observations, T0=None, nthreads=0):
    T0 = observations.shape[0] if T0 is None else min(T0, observations.shape[0])
    Y = np.hstack([*observations[:T0]])
    return np.stack([
        dense_logpdf_rows(Y, dense.obs_mean(T0, par.m[c], par.A[c], par.H[c]),
                          dense.obs_cov(T0, par.S[c], par.A[c], par.G[c], par.H[c], par.Lam[c]), nthreads)
        for c in range(par.n_clusters)])


def seq_loglik_matrix(packed, states, observations, T0=None, nthreads=0, raise_on_flag=False):
    """(n_c, n) log-likelihoods, sequential (prediction-error) formulation.
    ``states=None`` marginalises all hidden states (observation-only predict)."""
    observations = np.ascontiguousarray(observations, dtype=np.float64)
    T, n, l = observations.shape
    if states is not None:
        states = np.ascontiguousarray(states, dtype=np.float64)
        d = states.shape[2]
    packed = np.ascontiguousarray(packed, dtype=np.float64)
    n_c = packed.shape[0]
    if states is None:
        # stride = 1 + d + 3 d^2 + d l + l^2  ->  solve for d
        d = next(dd for dd in range(1, 17) if 1 + dd + 3 * dd * dd + dd * l + l * l == packed.shape[1])
    T0 = T if T0 is None else min(T0, T)
    ll = np.empty((n_c, n))
    st = lib().orc_seq_loglik(_p(states) if states is not None else None, _p(observations), T, n, d, l,
                              _p(packed), n_c, T0, _p(ll), nthreads)
    if st < 0:
        raise ValueError("dimension too large for the C oracle")
    if st and raise_on_flag:
        raise np.linalg.LinAlgError("Array must not contain infs or NaNs.")
    return ll


def assign(ll, pi):
    ll = np.ascontiguousarray(ll, dtype=np.float64)
    pi = np.ascontiguousarray(pi, dtype=np.float64)
    out = np.empty(ll.shape[1], dtype=np.int64)
    lib().orc_assign(_p(ll), _p(pi), ll.shape[0], ll.shape[1], out.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
    return out


def e_step_dense(par, states, observations, old):
    new = dense.assign_from_loglik(par.pi, dense_loglik_matrix(par, states, observations))
    return new, int(np.sum(new != old))


def e_step_seq(par, states, observations, old):
    new = assign(seq_loglik_matrix(par.packed(), states, observations), par.pi)
    return new, int(np.sum(new != old))
