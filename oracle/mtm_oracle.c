/*
 * CPU oracle (plain C) for the hard-assignment EM hot path of
 * burkh4rt/Multimodal-Trajectory-Modeling.
 *
 * TEST INFRASTRUCTURE ONLY: linked/loaded by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.  Never by the product package.
 *
 * Two independent restatements of the per-(trajectory, cluster) log-likelihood:
 *
 *  orc_dense_loglik   the reference's own formulation: for every row, sub-set the dense
 *                     joint covariance to the coordinates where (y - mean) is finite, take
 *                     det and solve by LU with partial pivoting (what numba lowers
 *                     np.linalg.det / np.linalg.solve to: LAPACK getrf / gesv), and form
 *                     -1/2 log((2 pi)^k det) - 1/2 r' Sigma^-1 r.
 *                     Follows framework/marginalizable_state_space_model.py:728-773.
 *                     Two LU factorisations per row, like the reference.
 *
 *  orc_seq_loglik     the prediction-error (sequential conditioning) form of the same
 *                     density, licensed by the identity the reference tests at
 *                     marginalizable_state_space_model.py:993-1002; mean/covariance
 *                     recursions from :62-64, :100-105, :161-166, :217-222, :312-339.
 *                     O(T) instead of O(T^3); used for parity at 10^5..10^7 trajectories.
 *
 * Parity status: PINNED through tests/test_oracle_golden.py (both functions are checked
 * against tests/golden (npz files), which hold outputs of the unmodified reference).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAXD 16

static int is_finite(double v) { return isfinite(v); }

/* LU with partial pivoting in place (row-major k x k); returns sign of the permutation,
 * 0 if an exactly zero pivot is met (singular). */
static int lu_factor(double *a, int k, int *piv)
{
    int sign = 1;
    for (int c = 0; c < k; ++c) {
        int p = c;
        double best = fabs(a[c * k + c]);
        for (int r = c + 1; r < k; ++r) {
            double v = fabs(a[r * k + c]);
            if (v > best) { best = v; p = r; }
        }
        piv[c] = p;
        if (p != c) {
            for (int j = 0; j < k; ++j) {
                double t = a[c * k + j]; a[c * k + j] = a[p * k + j]; a[p * k + j] = t;
            }
            sign = -sign;
        }
        double d = a[c * k + c];
        if (d == 0.0) return 0;
        for (int r = c + 1; r < k; ++r) {
            double f = a[r * k + c] / d;
            a[r * k + c] = f;
            for (int j = c + 1; j < k; ++j) a[r * k + j] -= f * a[c * k + j];
        }
    }
    return sign;
}

static void lu_solve(const double *a, int k, const int *piv, double *b)
{
    for (int c = 0; c < k; ++c) {
        int p = piv[c];
        if (p != c) { double t = b[c]; b[c] = b[p]; b[p] = t; }
    }
    for (int r = 1; r < k; ++r) {
        double s = b[r];
        for (int j = 0; j < r; ++j) s -= a[r * k + j] * b[j];
        b[r] = s;
    }
    for (int r = k - 1; r >= 0; --r) {
        double s = b[r];
        for (int j = r + 1; j < k; ++j) s -= a[r * k + j] * b[j];
        b[r] = s / a[r * k + r];
    }
}

/* return: 0 ok; 1 = a used covariance sub-block holds NaN/inf (the reference raises
 * LinAlgError there); 2 = exactly singular sub-block (the reference raises). */
int orc_dense_loglik(const double *Y, long n, int K, const double *mean, const double *cov,
                     double *out, int nthreads)
{
    int status = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        double *sub = (double *)malloc(sizeof(double) * (size_t)K * K);
        double *sub2 = (double *)malloc(sizeof(double) * (size_t)K * K);
        double *r = (double *)malloc(sizeof(double) * K);
        double *sol = (double *)malloc(sizeof(double) * K);
        int *idx = (int *)malloc(sizeof(int) * K);
        int *piv = (int *)malloc(sizeof(int) * K);
#pragma omp for schedule(static)
        for (long i = 0; i < n; ++i) {
            int k = 0;
            for (int j = 0; j < K; ++j) {
                double v = Y[i * K + j] - mean[j];
                if (is_finite(v)) { idx[k] = j; r[k] = v; ++k; }
            }
            if (k == 0) { out[i] = -0.0; continue; }
            int bad = 0;
            for (int a = 0; a < k; ++a)
                for (int b = 0; b < k; ++b) {
                    double v = cov[(size_t)idx[a] * K + idx[b]];
                    if (!is_finite(v)) bad = 1;
                    sub[a * k + b] = v;
                    sub2[a * k + b] = v;
                }
            if (bad) {
#pragma omp atomic write
                status = 1;
                out[i] = NAN;
                continue;
            }
            /* first factorisation: determinant */
            int sign = lu_factor(sub, k, piv);
            double det = (double)sign;
            for (int a = 0; a < k; ++a) det *= sub[a * k + a];
            /* second factorisation: solve */
            double q = 0.0;
            int sign2 = lu_factor(sub2, k, piv);
            if (sign2 == 0) {
#pragma omp atomic write
                status = 2;
                out[i] = NAN;
                continue;
            }
            for (int a = 0; a < k; ++a) sol[a] = r[a];
            lu_solve(sub2, k, piv, sol);
            for (int a = 0; a < k; ++a) q += r[a] * sol[a];
            out[i] = -0.5 * log(pow(2.0 * M_PI, (double)k) * det) - 0.5 * q;
        }
        free(sub); free(sub2); free(r); free(sol); free(idx); free(piv);
    }
    return status;
}

/* ------------------------------------------------------------------------------------ */
/* sequential (prediction-error) form                                                   */
/* ------------------------------------------------------------------------------------ */

/* packed parameter row: [pi, m(d), S(d*d), A(d*d), G(d*d), H(d*l), Lam(l*l)], row-major */
static long par_stride(int d, int l) { return 1 + d + 3 * d * d + d * l + l * l; }

/* One (trajectory, cluster) evaluation.  zt/xt: pointers to step-0 values of this trajectory,
 * step strides zs/xs (in doubles).  use_z = 0: hidden states are not read (all missing).
 * Returns the log-likelihood; *flag |= 1 if a pivot was non-finite. */
static double seq_eval(const double *zt, long zs, const double *xt, long xs, int T0, int d, int l,
                       const double *par, int use_z, int *flag)
{
    const double *m = par + 1, *S = m + d, *A = S + d * d, *G = A + d * d, *H = G + d * d,
                 *Lam = H + d * l;
    double mu[ORC_MAXD], P[ORC_MAXD][ORC_MAXD], mu0[ORC_MAXD];
    double C[ORC_MAXD][ORC_MAXD], Sx[ORC_MAXD][ORC_MAXD], mux[ORC_MAXD], mux0[ORC_MAXD];
    double tmp[ORC_MAXD][ORC_MAXD], tv[ORC_MAXD];
    for (int i = 0; i < d; ++i) {
        mu[i] = m[i]; mu0[i] = m[i];
        for (int j = 0; j < d; ++j) P[i][j] = S[i * d + j];
    }
    double q = 0.0, mant = 1.0;
    long expo = 0;
    int k = 0;
    for (int t = 0; t < T0; ++t) {
        if (t > 0) {
            /* mu <- mu A ; mu0 <- mu0 A (unconditional mean, decides which coordinates count) */
            for (int j = 0; j < d; ++j) {
                double s = 0.0, s0 = 0.0;
                for (int i = 0; i < d; ++i) { s += mu[i] * A[i * d + j]; s0 += mu0[i] * A[i * d + j]; }
                tv[j] = s; tmp[0][j] = s0;
            }
            for (int j = 0; j < d; ++j) { mu[j] = tv[j]; mu0[j] = tmp[0][j]; }
            /* P <- A' P A + G */
            for (int i = 0; i < d; ++i)
                for (int j = 0; j < d; ++j) {
                    double s = 0.0;
                    for (int a = 0; a < d; ++a) s += P[i][a] * A[a * d + j];
                    tmp[i][j] = s;
                }
            for (int i = 0; i < d; ++i)
                for (int j = 0; j < d; ++j) {
                    double s = G[i * d + j];
                    for (int a = 0; a < d; ++a) s += A[a * d + i] * tmp[a][j];
                    P[i][j] = s;
                }
        }
        /* hidden-state coordinates that are present */
        if (use_z) {
            for (int j = 0; j < d; ++j) {
                double y = zt[t * zs + j];
                if (!is_finite(y - mu0[j])) continue;
                double s = P[j][j], v = y - mu[j];
                if (!is_finite(s)) *flag |= 1;
                q += v * v / s;
                mant *= s; ++k;
                double col[ORC_MAXD];
                for (int i = 0; i < d; ++i) col[i] = P[i][j];
                for (int i = 0; i < d; ++i) {
                    double g = col[i] / s;
                    mu[i] += g * v;
                    for (int b = 0; b < d; ++b) P[i][b] -= g * col[b];
                }
                for (int i = 0; i < d; ++i) { P[i][j] = 0.0; P[j][i] = 0.0; }
                mu[j] = y;
            }
        }
        /* predictive of x_t given everything conditioned on so far */
        for (int b = 0; b < l; ++b) {
            double s = 0.0, s0 = 0.0;
            for (int i = 0; i < d; ++i) { s += mu[i] * H[i * l + b]; s0 += mu0[i] * H[i * l + b]; }
            mux[b] = s; mux0[b] = s0;
        }
        for (int i = 0; i < d; ++i)
            for (int b = 0; b < l; ++b) {
                double s = 0.0;
                for (int a = 0; a < d; ++a) s += P[i][a] * H[a * l + b];
                C[i][b] = s;
            }
        for (int a = 0; a < l; ++a)
            for (int b = 0; b < l; ++b) {
                double s = Lam[a * l + b];
                for (int i = 0; i < d; ++i) s += H[i * l + a] * C[i][b];
                Sx[a][b] = s;
            }
        for (int j = 0; j < l; ++j) {
            double y = xt[t * xs + j];
            if (!is_finite(y - mux0[j])) continue;
            double s = Sx[j][j], v = y - mux[j];
            if (!is_finite(s)) *flag |= 1;
            q += v * v / s;
            mant *= s; ++k;
            double cz[ORC_MAXD], cx[ORC_MAXD];
            for (int i = 0; i < d; ++i) cz[i] = C[i][j];
            for (int b = 0; b < l; ++b) cx[b] = Sx[j][b];
            for (int i = 0; i < d; ++i) {
                double g = cz[i] / s;
                mu[i] += g * v;
                for (int a = 0; a < d; ++a) P[i][a] -= g * cz[a];
                for (int b = 0; b < l; ++b) C[i][b] -= g * cx[b];
            }
            for (int a = 0; a < l; ++a) {
                double g = cx[a] / s;
                mux[a] += g * v;
                for (int b = 0; b < l; ++b) Sx[a][b] -= g * cx[b];
            }
        }
        /* keep the running determinant in range */
        int e;
        mant = frexp(mant, &e);
        expo += e;
    }
    if (k == 0) return -0.0;
    double logdet = log(mant) + (double)expo * M_LN2;
    return -0.5 * ((double)k * log(2.0 * M_PI) + logdet) - 0.5 * q;
}

/* z: (T, n, d), x: (T, n, l) C-contiguous; params: (n_c, stride); ll: (n_c, n).
 * z may be NULL (hidden states marginalised).  Returns OR of flags. */
int orc_seq_loglik(const double *z, const double *x, int T, long n, int d, int l,
                   const double *params, int n_c, int T0, double *ll, int nthreads)
{
    if (d > ORC_MAXD || l > ORC_MAXD) return -1;
    if (T0 > T) T0 = T;
    int status = 0;
    long stride = par_stride(d, l);
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static) reduction(| : status)
    for (long i = 0; i < n; ++i) {
        for (int c = 0; c < n_c; ++c) {
            int flag = 0;
            ll[(size_t)c * n + i] =
                seq_eval(z ? z + i * d : NULL, (long)n * d, x + i * l, (long)n * l, T0, d, l,
                         params + c * stride, z != NULL, &flag);
            status |= flag;
        }
    }
    return status;
}

/* linear-domain argmax of pi_c * exp(ll_c): first maximum wins, first NaN wins
 * (np.argmax semantics; framework/marginalizable_mixture_model.py:760-771) */
void orc_assign(const double *ll, const double *pi, int n_c, long n, int64_t *assign)
{
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        double best = pi[0] * exp(ll[i]);
        int arg = 0;
        for (int c = 1; c < n_c; ++c) {
            if (isnan(best)) break;
            double v = pi[c] * exp(ll[(size_t)c * n + i]);
            if (v > best || isnan(v)) { best = v; arg = c; }
        }
        assign[i] = arg;
    }
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
