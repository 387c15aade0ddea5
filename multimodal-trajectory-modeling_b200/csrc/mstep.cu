// Kernel 2: M-step of the hard-assignment EM as per-cluster sufficient statistics + a tiny finalise.
//
// The reference's M_step (framework/marginalizable_mixture_model.py:1254-1308) gathers, per
// cluster, the rows that are completely finite and calls regress / regress_alpha (:1167-1233)
// and np.mean / np.cov.  All of that is a function of these moments of the member set I_c:
//   N_c                                                            -> pi_c            (:1261-1263)
//   n0, sum z_1, sum z_1 z_1'      over i with finite z_1          -> m_c, S_c        (:1267-1270)
//   n1, sum u, sum u u'  u=[z_t, z_{t+1}] over (i,t) both finite   -> A_c, Gamma_c    (:1272-1290)
//   n2, sum w, sum w w'  w=[z_t, x_t]    over (i,t) both finite   -> H_c, Lambda_c   (:1292-1308)
//
// stats kernel: one thread per trajectory accumulates its own moments in registers (two passes
// over its T steps: {init, transition} then {measurement}); the block then reduces them BY
// CLUSTER: a stable counting sort of the tile's 128 labels gives each cluster its member list in
// ascending trajectory order, the per-trajectory moments go through shared memory (transposed),
// and thread (c, f) adds feature f of cluster c's members in list order — no floating-point
// atomics anywhere, the summation order is a pure function of (n, grid).  Block sums are carried
// as double-double (hi, lo) pairs across tiles and blocks, so the accumulated rounding error does
// not grow with n.
#include <cstdlib>

#include "common.cuh"
#include "filter.cuh"
#include "stats_tile.cuh"

namespace mtm {

template <int D_, int L_>
__global__ void __launch_bounds__(kSThreads) stats_kernel(const StatsArgs a) {
    using DM = Dims<D_, L_>;
    const DM dm{a.d, a.l};
    const int r = blockIdx.y;  // batched restarts
    if (a.check_done && a.ctrl[r].done) return;
    const int32_t *const assign = a.assign + (size_t)r * a.n;
    const int n_c = a.n_c;
    const int sstride = stats_stride(dm.d(), dm.l());
    extern __shared__ __align__(16) unsigned char smem_raw[];
    StatsSmem sm;
    sm.carve(smem_raw, n_c, dm.d(), dm.l());
    const int tid = threadIdx.x;
    for (int k = tid; k < 2 * n_c * sstride; k += kSThreads) sm.acc[k] = 0.0;
    __syncthreads();
    const int64_t n = a.n;
    const int64_t n_tiles = (n + kSThreads - 1) / kSThreads;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t i_raw = tile * kSThreads + tid;
        const bool valid = i_raw < n;
        const int64_t i = valid ? i_raw : n - 1;
        stats_tile<DM, true>(dm, a.z, a.x, a.ld, a.T, i, valid ? assign[i] : -1, n_c, smem_raw, a.pivot);
    }
    __syncthreads();
    double *out = a.partials + 2 * ((size_t)r * gridDim.x + blockIdx.x) * n_c * sstride;
    for (int k = tid; k < 2 * n_c * sstride; k += kSThreads) out[k] = sm.acc[k];
}

// Two-stage fixed-order reduction of the per-block (hi, lo) partials.
//  stage 1: slice s (blockIdx.y) adds blocks s, s + S, s + 2S, ... in that order  -> scratch[s][v]
//  stage 2: thread v adds the S slices in order                                  -> out[v]
// tail slots (kTail, common.cuh): out[n_vals] = n_switches of the preceding E-step (exact integer as double),
//             out[n_vals+1] = local n, out[n_vals+2..3] = this rank's NONFINITE / SINGULAR flags (0 or 1)
__global__ void stats_reduce1_kernel(const double *__restrict__ partials_all, int grid, int n_vals,
                                     double *scratch_all, const Ctrl *ctrl, int check_done, const StatsCache *cache,
                                     int grid_delta) {
    const int r = blockIdx.z;  // batched restarts
    if (check_done && ctrl[r].done) return;
    if (cache && cache->mode) grid = grid_delta;  // the delta twin wrote the partials: one wave of blocks
    const double *partials = partials_all + 2 * (size_t)r * grid * n_vals;
    double *scratch = scratch_all + 2 * (size_t)r * gridDim.y * n_vals;
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y, S = gridDim.y;
    if (v >= n_vals) return;
    double hi = 0.0, lo = 0.0;
    int b = s;
    for (; b + 3 * S < grid; b += 4 * S) {  // four independent loads in flight, added in block order
        const double2 p0 = __ldcs(reinterpret_cast<const double2 *>(partials + 2 * ((size_t)b * n_vals + v)));
        const double2 p1 = __ldcs(reinterpret_cast<const double2 *>(partials + 2 * ((size_t)(b + S) * n_vals + v)));
        const double2 p2 = __ldcs(reinterpret_cast<const double2 *>(partials + 2 * ((size_t)(b + 2 * S) * n_vals + v)));
        const double2 p3 = __ldcs(reinterpret_cast<const double2 *>(partials + 2 * ((size_t)(b + 3 * S) * n_vals + v)));
        dd_add(hi, lo, p0.x);
        lo += p0.y;
        dd_add(hi, lo, p1.x);
        lo += p1.y;
        dd_add(hi, lo, p2.x);
        lo += p2.y;
        dd_add(hi, lo, p3.x);
        lo += p3.y;
    }
    for (; b < grid; b += S) {
        const double2 p = __ldcs(reinterpret_cast<const double2 *>(partials + 2 * ((size_t)b * n_vals + v)));
        dd_add(hi, lo, p.x);
        lo += p.y;
    }
    *reinterpret_cast<double2 *>(scratch + 2 * ((size_t)s * n_vals + v)) = make_double2(hi, lo);
}

__global__ void stats_reduce2_kernel(const double *__restrict__ scratch_all, int S, int n_vals, double *out_all,
                                     const Ctrl *ctrl, int check_done, int with_switches, double n_local,
                                     const StatsCache *cache, double *bases) {
    const int r = blockIdx.y;
    if (check_done && ctrl[r].done) return;
    const double *scratch = scratch_all + 2 * (size_t)r * S * n_vals;
    double *out = out_all + (size_t)r * (n_vals + kTail);
    const unsigned long long *n_switches_src = with_switches ? &ctrl[r].n_switches : nullptr;
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < n_vals) {
        double hi = 0.0, lo = 0.0;
        for (int s = 0; s < S; ++s) {
            const double2 p = *reinterpret_cast<const double2 *>(scratch + 2 * ((size_t)s * n_vals + v));
            dd_add(hi, lo, p.x);
            lo += p.y;
        }
        if (cache) {  // statistics cache (single fit): keep the sums of these labels as a (hi, lo) pair
            double *bd = bases + 2 * ((size_t)cache->dst * n_vals + v);
            if (cache->mode) {  // the partials were differences against entry src
                const double *bs = bases + 2 * ((size_t)cache->src * n_vals + v);
                double bh = bs[0], bl = bs[1];
                dd_add(bh, bl, hi);
                bl += lo;
                hi = bh + bl;
                lo = bl - (hi - bh);
            }
            bd[0] = hi;
            bd[1] = lo;
        }
        out[v] = hi + lo;
    }
    if (v == 0) {
        const int st = ctrl[r].status;
        out[n_vals] = n_switches_src ? (double)(*n_switches_src) : 0.0;
        out[n_vals + 1] = n_local;
        out[n_vals + 2] = (st & MTM_STATUS_NONFINITE) ? 1.0 : 0.0;
        out[n_vals + 3] = (st & MTM_STATUS_SINGULAR) ? 1.0 : 0.0;
    }
}

// ------------------------------------------------------------------------------------------------
// finalise: parameters from moments.  One thread per cluster; run-time dimensions; tiny.
// ------------------------------------------------------------------------------------------------
constexpr int MX = MTM_MAX_DIM;

__device__ inline double qnan() { return __longlong_as_double(0x7ff8000000000000LL); }

// Compile-time (P_, Q_ > 0: everything unrolled, matrices in registers) or run-time sizes.
template <int P_, int Q_>
struct RegDims {
    static constexpr bool kStatic = (P_ > 0);
    static constexpr int PM = kStatic ? P_ : MX;
    static constexpr int QM = kStatic ? Q_ : MX;
    int p_rt, q_rt;
    __device__ __forceinline__ int p() const { return kStatic ? P_ : p_rt; }
    __device__ __forceinline__ int q() const { return kStatic ? Q_ : q_rt; }
};

// Cholesky factor of (M + alpha I) for SPD M (p x p), in place (lower triangle).  false if a pivot is not positive.
template <class RD>
__device__ __forceinline__ bool chol_factor(const RD &rd, double (&M)[RD::PM][RD::PM], double alpha) {
    const int p = rd.p();
    bool ok = true;
    MTM_UNROLL
    for (int j = 0; j < RD::PM; ++j) {
        if (j < p) {
            const double djj = M[j][j] + alpha;
            double s = djj;
            MTM_UNROLL
            for (int k = 0; k < RD::PM; ++k)
                if (k < j) s -= M[j][k] * M[j][k];
            // least squares: a pivot below ~64 eps of its diagonal entry is rounding noise of a rank-deficient
            // scatter matrix (exactly dependent columns leave +-1e-17 here), not information: hand over to the
            // minimum-norm solve.  Ridge: alpha keeps every pivot well above that; only non-finite input fails.
            ok = ok && (alpha > 0.0 ? s > 0.0 : s > 1.5e-14 * djj);
            const double dj = sqrt(s);
            M[j][j] = dj;
            const double inv = 1.0 / dj;
            MTM_UNROLL
            for (int i = 0; i < RD::PM; ++i) {
                if (i > j && i < p) {
                    double t = M[i][j];
                    MTM_UNROLL
                    for (int k = 0; k < RD::PM; ++k)
                        if (k < j) t -= M[i][k] * M[j][k];
                    M[i][j] = t * inv;
                }
            }
        }
    }
    return ok;
}

// v <- (L L')^-1 v with the factor from chol_factor
template <class RD>
__device__ __forceinline__ void chol_apply(const RD &rd, const double (&M)[RD::PM][RD::PM], double (&v)[RD::PM]) {
    const int p = rd.p();
    MTM_UNROLL
    for (int i = 0; i < RD::PM; ++i) {
        if (i < p) {
            double t = v[i];
            MTM_UNROLL
            for (int k = 0; k < RD::PM; ++k)
                if (k < i) t -= M[i][k] * v[k];
            v[i] = t / M[i][i];
        }
    }
    MTM_UNROLL
    for (int ii = 0; ii < RD::PM; ++ii) {
        const int i = RD::PM - 1 - ii;
        if (i < p) {
            double t = v[i];
            MTM_UNROLL
            for (int k = 0; k < RD::PM; ++k)
                if (k > i && k < p) t -= M[k][i] * v[k];
            v[i] = t / M[i][i];
        }
    }
}

// Minimum-norm least squares  B = pinv(G) R  for a symmetric positive SEMI-definite G (p x p, run-time p <= MX)
// by a cyclic Jacobi eigen-decomposition: what np.linalg.lstsq(X, Y, rcond=-1) (MMM:1194) returns for a
// rank-deficient design X with G = X'X, R = X'Y.  Eigenvalues below 1e-13 of the largest count as zero (they are
// rounding noise of the moment matrix; lstsq itself cuts singular values at machine precision).  Run-time indexed
// arrays in local memory: this is the rare fallback of a ~100-thread kernel.
__device__ __noinline__ void pinv_solve(int p, int q, double (*G)[MX], double (*R)[MX], double (*B)[MX]) {
    double V[MX][MX];
    for (int i = 0; i < p; ++i)
        for (int j = 0; j < p; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < p; ++i) {
            diag += G[i][i] * G[i][i];
            for (int j = i + 1; j < p; ++j) off += G[i][j] * G[i][j];
        }
        if (!(off > 1e-40 * diag)) break;
        for (int a = 0; a < p; ++a) {
            for (int b = a + 1; b < p; ++b) {
                const double g = G[a][b];
                if (g == 0.0) continue;
                const double theta = (G[b][b] - G[a][a]) / (2.0 * g);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
                for (int k = 0; k < p; ++k) {  // G <- G J (columns a, b)
                    const double ga = G[k][a], gb = G[k][b];
                    G[k][a] = c * ga - sn * gb;
                    G[k][b] = sn * ga + c * gb;
                }
                for (int k = 0; k < p; ++k) {  // G <- J' G (rows a, b)
                    const double ga = G[a][k], gb = G[b][k];
                    G[a][k] = c * ga - sn * gb;
                    G[b][k] = sn * ga + c * gb;
                }
                for (int k = 0; k < p; ++k) {
                    const double va = V[k][a], vb = V[k][b];
                    V[k][a] = c * va - sn * vb;
                    V[k][b] = sn * va + c * vb;
                }
            }
        }
    }
    double lmax = 0.0;
    for (int i = 0; i < p; ++i) lmax = fmax(lmax, G[i][i]);
    for (int i = 0; i < p; ++i)
        for (int c = 0; c < q; ++c) B[i][c] = 0.0;
    for (int k = 0; k < p; ++k) {
        const double lam = G[k][k];
        if (!(lam > 1e-13 * lmax)) continue;
        for (int c = 0; c < q; ++c) {
            double t = 0.0;
            for (int i = 0; i < p; ++i) t += V[i][k] * R[i][c];
            t /= lam;
            for (int i = 0; i < p; ++i) B[i][c] += V[i][k] * t;
        }
    }
}

// moments block [cnt | sum v (K) | upper(sum vv')] of v = [a (p) | b (q)], taken ABOUT THE PIVOT (pa, pb)  ->
//   B = argmin |b - a B|^2 + alpha |B|^2 (p x q, NO intercept), Cov = cov(b - a B) with mean removed and N-1
// (regress / regress_alpha, MMM:1167-1233).
//
// Everything is done on the scatter about the sample mean, C = sum (v - vbar)(v - vbar)', which the centred
// moments give without cancellation.  The no-intercept normal equations are a rank-one update of it,
//   X'X = C_aa + n mu_a mu_a',  X'Y = C_ab + n mu_a mu_b'      (mu = true means = pivot + centred mean),
// solved with Sherman-Morrison on the factor of C_aa + alpha I:
//   B0 = (C_aa + alpha I)^-1 C_ab,  w = (C_aa + alpha I)^-1 mu_a,  B = B0 + n / (1 + n mu_a'w) * w (mu_b - B0'mu_a)'.
// The raw moment matrix (condition ~ |mean|^2 / var worse) is never factorised.  Residual scatter:
//   sum (r - rbar)(r - rbar)' = C_bb - B'C_ab - C_ab'B + B'C_aa B.
// alpha = 0 and C_aa not positive definite (rank-deficient or underdetermined design): lstsq's minimum-norm
// solution through pinv_solve on the reconstructed X'X.
template <class RD>
__device__ __forceinline__ int regress_from_moments(const RD &rd, const double *blk, double alpha, const double *pa,
                                                    const double *pb, double *B_out, double *C_out) {
    const int p = rd.p(), q = rd.q();
    const int K = p + q;
    constexpr int KM = RD::PM + RD::QM;
    const double cnt = blk[0];
    const double *sv = blk + 1;
    const double *up = blk + 1 + K;
    int status = 0;
    if (!(cnt > 0.0)) {  // empty design: lstsq / Ridge raise in the reference
        for (int k = 0; k < p * q; ++k) B_out[k] = qnan();
        for (int k = 0; k < q * q; ++k) C_out[k] = qnan();
        return MTM_STATUS_EMPTY_FIT;
    }
    double cm[KM];  // centred means
    MTM_UNROLL
    for (int a = 0; a < KM; ++a)
        if (a < K) cm[a] = sv[a] / cnt;
    double full[KM][KM];  // scatter about the sample mean
    {
        int o = 0;
        MTM_UNROLL
        for (int a = 0; a < KM; ++a) {
            if (a < K) {
                MTM_UNROLL
                for (int b = a; b < KM; ++b)
                    if (b < K) {
                        const double v = fma(-cnt * cm[a], cm[b], up[o + (b - a)]);
                        full[a][b] = v;
                        full[b][a] = v;
                    }
                o += K - a;
            }
        }
    }
    double mua[RD::PM], mub[RD::QM];  // true means
    MTM_UNROLL
    for (int i = 0; i < RD::PM; ++i)
        if (i < p) mua[i] = pa[i] + cm[i];
    MTM_UNROLL
    for (int c = 0; c < RD::QM; ++c)
        if (c < q) mub[c] = pb[c] + cm[p + c];

    double M[RD::PM][RD::PM], B[RD::PM][RD::QM];
    MTM_UNROLL
    for (int i = 0; i < RD::PM; ++i) {
        if (i < p) {
            MTM_UNROLL
            for (int j = 0; j < RD::PM; ++j)
                if (j < p) M[i][j] = full[i][j];
        }
    }
    if (chol_factor(rd, M, alpha)) {
        double w[RD::PM];
        MTM_UNROLL
        for (int i = 0; i < RD::PM; ++i)
            if (i < p) w[i] = mua[i];
        chol_apply(rd, M, w);
        double s = 1.0;
        MTM_UNROLL
        for (int i = 0; i < RD::PM; ++i)
            if (i < p) s = fma(cnt * mua[i], w[i], s);
        const double f = cnt / s;
        MTM_UNROLL
        for (int c = 0; c < RD::QM; ++c) {
            if (c < q) {
                double col[RD::PM];
                MTM_UNROLL
                for (int i = 0; i < RD::PM; ++i)
                    if (i < p) col[i] = full[i][p + c];
                chol_apply(rd, M, col);  // column c of B0
                double e = mub[c];
                MTM_UNROLL
                for (int i = 0; i < RD::PM; ++i)
                    if (i < p) e = fma(-mua[i], col[i], e);
                MTM_UNROLL
                for (int i = 0; i < RD::PM; ++i)
                    if (i < p) B[i][c] = fma(f * w[i], e, col[i]);
            }
        }
    } else if (alpha == 0.0) {
        // rank-deficient / underdetermined design: minimum-norm solution of the raw normal equations
        double G[MX][MX], R[MX][MX], Bm[MX][MX];
        bool fin = true;
        for (int i = 0; i < p; ++i) {
            for (int j = 0; j < p; ++j) {
                G[i][j] = fma(cnt * mua[i], mua[j], full[i][j]);
                fin = fin && finite_bits(G[i][j]);
            }
            for (int c = 0; c < q; ++c) {
                R[i][c] = fma(cnt * mua[i], mub[c], full[i][p + c]);
                fin = fin && finite_bits(R[i][c]);
            }
        }
        if (fin) {
            pinv_solve(p, q, G, R, Bm);
            MTM_UNROLL
            for (int i = 0; i < RD::PM; ++i)
                MTM_UNROLL
                for (int c = 0; c < RD::QM; ++c)
                    if (i < p && c < q) B[i][c] = Bm[i][c];
        } else {
            status |= MTM_STATUS_CHOL_FAIL;
            MTM_UNROLL
            for (int i = 0; i < RD::PM; ++i)
                MTM_UNROLL
                for (int c = 0; c < RD::QM; ++c) B[i][c] = qnan();
        }
    } else {
        status |= MTM_STATUS_CHOL_FAIL;  // ridge system not positive definite: non-finite moments
        MTM_UNROLL
        for (int i = 0; i < RD::PM; ++i)
            MTM_UNROLL
            for (int c = 0; c < RD::QM; ++c) B[i][c] = qnan();
    }
    // E = C_ab - C_aa B (zero for least squares up to rounding)
    double E[RD::PM][RD::QM];
    MTM_UNROLL
    for (int i = 0; i < RD::PM; ++i) {
        if (i < p) {
            MTM_UNROLL
            for (int c = 0; c < RD::QM; ++c) {
                if (c < q) {
                    double s = full[i][p + c];
                    MTM_UNROLL
                    for (int k = 0; k < RD::PM; ++k)
                        if (k < p) s -= full[i][k] * B[k][c];
                    E[i][c] = s;
                }
            }
        }
    }
    MTM_UNROLL
    for (int a = 0; a < RD::QM; ++a) {
        if (a < q) {
            MTM_UNROLL
            for (int b = a; b < RD::QM; ++b) {
                if (b < q) {
                    // sum (r - rbar)(r - rbar)' = C_bb - B'C_ab - E'B
                    double s = full[p + a][p + b];
                    MTM_UNROLL
                    for (int i = 0; i < RD::PM; ++i)
                        if (i < p) s -= B[i][a] * full[i][p + b];
                    MTM_UNROLL
                    for (int i = 0; i < RD::PM; ++i)
                        if (i < p) s -= E[i][a] * B[i][b];
                    const double v = (cnt > 1.0) ? s / (cnt - 1.0) : qnan();
                    C_out[a * q + b] = v;
                    C_out[b * q + a] = v;
                }
            }
        }
    }
    MTM_UNROLL
    for (int i = 0; i < RD::PM; ++i) {
        if (i < p) {
            MTM_UNROLL
            for (int c = 0; c < RD::QM; ++c)
                if (c < q) B_out[i * q + c] = B[i][c];
        }
    }
    return status;
}

template <int D_, int L_>
__global__ void finalize_kernel(const FinalizeArgs a0, const double *tail0) {
    const int r = blockIdx.x;  // batched restarts: one block per fit
    FinalizeArgs a = a0;
    Ctrl *ctrl = a0.ctrl + r;
    if (a.mode == 1 && ctrl->done) return;
    const int d = D_ > 0 ? D_ : a.d, l = D_ > 0 ? L_ : a.l, n_c = a.n_c;
    const int F1 = 1 + d + d * (d + 1) / 2, F2 = 1 + 2 * d + d * (2 * d + 1);
    const int sstride = stats_stride(d, l), pstride = param_stride(d, l);
    a.stats = a0.stats + (size_t)r * ((size_t)n_c * sstride + (tail0 ? kTail : 0));
    a.params = a0.params + (size_t)r * n_c * pstride;
    a.switch_hist = a0.switch_hist ? a0.switch_hist + (size_t)r * a.n_steps : nullptr;
    const double *tail = tail0 ? a.stats + (size_t)n_c * sstride : nullptr;
    __shared__ int s_go;
    __shared__ int s_status;
    if (threadIdx.x == 0) {
        double mn = 1e300;
        for (int c = 0; c < n_c; ++c) mn = fmin(mn, a.stats[(size_t)c * sstride]);
        int go = 1;
        s_status = 0;
        if (tail) {
            // failure bits of the preceding E-step, summed over the ranks: every rank adopts the union, so that
            // the decisions below (and the status the host finally sees) are the same everywhere
            const int st = (tail[2] > 0.0 ? MTM_STATUS_NONFINITE : 0) | (tail[3] > 0.0 ? MTM_STATUS_SINGULAR : 0);
            if (st) ctrl->status |= st;
        }
        if (a.mode == 1) {
            const long long sw = (long long)(tail ? tail[0] : 0.0);
            ctrl->min_count = (long long)mn;
            if (ctrl->n_m_steps == 0) {
                // before the first M-step: only the near-empty-cluster guard (MMM:1328-1336); the extended
                // mixture has none (its caller re-draws such an initial assignment, SSMM:251-260)
                if (a.variant == 0 && mn <= 3.0) {
                    ctrl->done = 1;
                    ctrl->exit_code = MTM_EXIT_GUARD0;
                    go = 0;
                }
            } else if (ctrl->status & (MTM_STATUS_NONFINITE | MTM_STATUS_SINGULAR | MTM_STATUS_COMM_TIMEOUT)) {
                // the reference's E-step raises LinAlgError here (det/solve on a non-finite or
                // singular covariance block): stop, the host re-raises.  A failed exchange (a peer never
                // arrived) leaves un-reduced statistics behind: stop as well instead of iterating on them.
                ctrl->done = 1;
                ctrl->exit_code = MTM_EXIT_ERROR;
                go = 0;
            } else if (a.variant == 1 && mn <= 0.0) {
                // extended mixture: its _E_step asserts that every cluster has a member BEFORE it stores the
                // new labels or counts switches (SSMM:122-124); the exception ends the fit (SSMM:275-276)
                ctrl->n_e_steps += 1;
                ctrl->done = 1;
                ctrl->exit_code = MTM_EXIT_EMPTY;
                go = 0;
            } else {
                if (a.switch_hist && ctrl->n_e_steps < a.n_steps) a.switch_hist[ctrl->n_e_steps] = sw;
                ctrl->n_e_steps += 1;
                ctrl->last_switches = sw;
                ctrl->n_switches = 0ull;
                if (sw == 0) {  // converged (MMM:1342-1345): parameters stay
                    ctrl->done = 1;
                    ctrl->exit_code = MTM_EXIT_CONVERGED;
                    go = 0;
                } else if (mn <= 3.0) {  // guard (MMM:1346-1356): parameters of the previous M-step stay
                    ctrl->done = 1;
                    ctrl->exit_code = MTM_EXIT_GUARD;
                    go = 0;
                }
            }
        } else {
            ctrl->min_count = (long long)mn;
        }
        s_go = go;
    }
    __syncthreads();
    if (!s_go) return;

    // one thread per (cluster, block of the parameter row): 0 = pi, m, S; 1 = A, Gamma; 2 = H, Lambda
    for (int task = threadIdx.x; task < 3 * n_c; task += blockDim.x) {
        const int c = task / 3, part = task - 3 * c;
        const double *st = a.stats + (size_t)c * sstride;
        double *pr = a.params + (size_t)c * pstride;
        const double n_total = tail ? tail[1] : (double)a.n_total;
        // util.regress (the extended components' least squares) carries a fixed 1e-6 ridge (util_state_space.py:33-61)
        const double alpha_eff = (a.variant == 1 && !(a.alpha > 0.0)) ? 1e-6 : a.alpha;
        int status = 0;
        if (part == 0) {
            pr[0] = st[0] / n_total;  // pi_c = mean(assign == c)
            // initial state: mean and np.cov(ddof=1) of the finite z_1 rows
            const double *b1 = st + 1;
            const double n0 = b1[0];
            double m[MX];  // mean of the centred values; the pivot goes back on for the parameter
            for (int i = 0; i < d; ++i) {
                m[i] = b1[1 + i] / n0;  // 0/0 = NaN for an empty sample, like np.mean
                pr[1 + i] = a0.pivot[i] + m[i];
            }
            if (a.variant == 1) {  // np.nanmean per coordinate (SSLG:43-44); the covariance below is unchanged
                const double *nm = a0.nanmean + ((size_t)r * n_c + c) * 2 * d;
                for (int i = 0; i < d; ++i) pr[1 + i] = nm[i] / nm[d + i];
            }
            int o = 1 + d;
            double *S = pr + 1 + d;
            for (int i = 0; i < d; ++i)
                for (int j = i; j < d; ++j) {
                    const double v = (n0 > 1.0) ? fma(-n0 * m[i], m[j], b1[o]) / (n0 - 1.0) : qnan();
                    S[i * d + j] = v;
                    S[j * d + i] = v;
                    ++o;
                }
        } else if (part == 1) {
            double *A = pr + 1 + d + d * d, *G = A + d * d;
            const RegDims<D_, D_> rd{d, d};
            status |= regress_from_moments(rd, st + 1 + F1, alpha_eff, a0.pivot, a0.pivot, A, G);
        } else {
            double *H = pr + 1 + d + 3 * d * d, *Lam = H + d * l;
            const RegDims<D_, L_> rd{d, l};
            status |= regress_from_moments(rd, st + 1 + F1 + F2, alpha_eff, a0.pivot, a0.pivot + MX, H, Lam);
        }
        if (status) atomicOr(&s_status, status);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_status) atomicOr(&ctrl->status, s_status);
        ctrl->n_m_steps += 1;
        if (a.mode == 1 && s_status) {  // lstsq / Ridge raise on an empty design in the reference
            ctrl->done = 1;
            ctrl->exit_code = MTM_EXIT_ERROR;
        } else if (a.mode == 1 && ctrl->n_e_steps >= a.n_steps) {  // all n_steps E-steps used (MMM:1340)
            ctrl->done = 1;
            ctrl->exit_code = MTM_EXIT_EXHAUSTED;
        }
    }
}

// ---- pivot: per-feature sum and count of the finite entries of a (T, n, w) array -------------------------------
// Fixed-order reductions (warp shuffles, then warps, then blocks): the same bits every run.
__global__ void __launch_bounds__(256) pivot_partial_kernel(const double *__restrict__ a, int64_t n, int64_t ld, int T,
                                                            int w, double *partials) {
    double s[MX], c[MX];
    MTM_UNROLL
    for (int j = 0; j < MX; ++j) s[j] = c[j] = 0.0;
    const int64_t rows = (int64_t)T * n;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = r / n, i = r - t * n;
        const double *row = a + (t * ld + i) * w;
        MTM_UNROLL
        for (int j = 0; j < MX; ++j) {
            if (j < w) {
                const double y = __ldg(row + j);
                const bool f = finite_bits(y);
                s[j] += f ? y : 0.0;
                c[j] += f ? 1.0 : 0.0;
            }
        }
    }
    __shared__ double sh[8][2 * MX];
    MTM_UNROLL
    for (int j = 0; j < MX; ++j) {
        for (int o = 16; o > 0; o >>= 1) {
            s[j] += __shfl_xor_sync(0xffffffffu, s[j], o);
            c[j] += __shfl_xor_sync(0xffffffffu, c[j], o);
        }
        if ((threadIdx.x & 31) == 0) {
            sh[threadIdx.x >> 5][j] = s[j];
            sh[threadIdx.x >> 5][MX + j] = c[j];
        }
    }
    __syncthreads();
    if (threadIdx.x < 2 * MX) {
        double t = 0.0;
        for (int k = 0; k < 8; ++k) t += sh[k][threadIdx.x];
        partials[(size_t)blockIdx.x * 2 * MX + threadIdx.x] = t;
    }
}

__global__ void pivot_reduce_kernel(const double *__restrict__ partials, int grid, double *out16) {
    if (threadIdx.x < 2 * MX) {
        double t = 0.0;
        for (int b = 0; b < grid; ++b) t += partials[(size_t)b * 2 * MX + threadIdx.x];
        out16[threadIdx.x] = t;
    }
}

size_t pivot_scratch_doubles(int sm_count) { return (size_t)sm_count * 8 * 2 * MX; }

cudaError_t launch_pivot_sums(const double *a, int64_t n, int64_t ld, int T, int w, int sm_count, double *scratch,
                              double *out16, cudaStream_t s) {
    const int64_t rows = (int64_t)T * n;
    int grid = (int)((rows + 255) / 256);
    if (grid > sm_count * 8) grid = sm_count * 8;
    if (grid < 1) grid = 1;
    pivot_partial_kernel<<<grid, 256, 0, s>>>(a, n, ld, T, w, scratch);
    pivot_reduce_kernel<<<1, 32, 0, s>>>(scratch, grid, out16);
    return cudaGetLastError();
}

// ---- framework_extended helpers ---------------------------------------------------------------------------------
// np.nanmean of z_1 per cluster and coordinate as [sum | count] of the non-NaN entries (an inf counts, like numpy).
// One warp per 32 trajectories; per cluster present in the warp a fixed-order shuffle reduction, lane 0 adds into
// the warp's shared-memory row; warps, then blocks, are added in index order: the same bits every run.
__global__ void __launch_bounds__(256) nanmean_partial_kernel(const double *__restrict__ z, int64_t n, int d,
                                                              const int32_t *__restrict__ assign, int n_c,
                                                              double *partials, const Ctrl *ctrl, int check_done) {
    if (check_done && ctrl->done) return;
    extern __shared__ double s_nm[];  // [8 warps][n_c][2d]
    const int w2 = 2 * d, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = threadIdx.x; k < 8 * n_c * w2; k += 256) s_nm[k] = 0.0;
    __syncthreads();
    double *mine = s_nm + (size_t)warp * n_c * w2;
    const int64_t n_warp_tiles = (n + 31) / 32;
    for (int64_t wt = (int64_t)blockIdx.x * 8 + warp; wt < n_warp_tiles; wt += (int64_t)gridDim.x * 8) {
        const int64_t i = wt * 32 + lane;
        const int lab = i < n ? assign[i] : -1;
        double v[MX];
        MTM_UNROLL
        for (int j = 0; j < MX; ++j) v[j] = (j < d && i < n) ? __ldg(z + i * d + j) : qnan();
        for (int c = 0; c < n_c; ++c) {
            if (__ballot_sync(0xffffffffu, lab == c) == 0u) continue;  // warp-uniform
            MTM_UNROLL
            for (int j = 0; j < MX; ++j) {
                if (j < d) {
                    const bool use = (lab == c) && (v[j] == v[j]);
                    double sm = use ? v[j] : 0.0, ct = use ? 1.0 : 0.0;
                    for (int o = 16; o > 0; o >>= 1) {
                        sm += __shfl_xor_sync(0xffffffffu, sm, o);
                        ct += __shfl_xor_sync(0xffffffffu, ct, o);
                    }
                    if (lane == 0) {
                        mine[c * w2 + j] += sm;
                        mine[c * w2 + d + j] += ct;
                    }
                }
            }
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n_c * w2; k += 256) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += s_nm[(size_t)w * n_c * w2 + k];
        partials[(size_t)blockIdx.x * n_c * w2 + k] = t;
    }
}

__global__ void nanmean_reduce_kernel(const double *__restrict__ partials, int grid, int n_vals, double *out,
                                      const Ctrl *ctrl, int check_done) {
    if (check_done && ctrl->done) return;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_vals) return;
    double t = 0.0;
    for (int b = 0; b < grid; ++b) t += partials[(size_t)b * n_vals + k];
    out[k] = t;
}

static int nanmean_grid(int64_t n, int sm_count) {
    const int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)sm_count * 4;
    return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

size_t nanmean_scratch_doubles(int n_c, int d, int sm_count) { return (size_t)sm_count * 4 * n_c * 2 * d; }

cudaError_t launch_nanmean(const double *z, int64_t n, int d, const int32_t *assign, int n_c, int sm_count,
                           double *scratch, double *out, const Ctrl *ctrl, int check_done, cudaStream_t s) {
    const int grid = nanmean_grid(n, sm_count);
    const int n_vals = n_c * 2 * d;
    const size_t smem = (size_t)8 * n_vals * sizeof(double);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(nanmean_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    nanmean_partial_kernel<<<grid, 256, smem, s>>>(z, n, d, assign, n_c, scratch, ctrl, check_done);
    nanmean_reduce_kernel<<<(n_vals + 127) / 128, 128, 0, s>>>(scratch, grid, n_vals, out, ctrl, check_done);
    return cudaGetLastError();
}

// labels before an E-step of the extended fit (kept so that a discarded E-step can be undone); honours `done`
__global__ void copy_labels_kernel(int32_t *__restrict__ dst, const int32_t *__restrict__ src, int64_t n,
                                   const Ctrl *ctrl, int check_done) {
    if (check_done && ctrl->done) return;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

cudaError_t launch_copy_labels(int32_t *dst, const int32_t *src, int64_t n, const Ctrl *ctrl, int check_done,
                               cudaStream_t s) {
    int grid = (int)((n + 255) / 256);
    if (grid > 1184) grid = 1184;
    if (grid < 1) grid = 1;
    copy_labels_kernel<<<grid, 256, 0, s>>>(dst, src, n, ctrl, check_done);
    return cudaGetLastError();
}

__global__ void counts_kernel(const int32_t *__restrict__ assign, int64_t n, int n_c, unsigned long long *counts) {
    __shared__ unsigned int s[MTM_MAX_CLUSTERS];
    for (int c = threadIdx.x; c < n_c; c += blockDim.x) s[c] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int a = assign[i];
        if (a >= 0 && a < n_c) atomicAdd(&s[a], 1u);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < n_c; c += blockDim.x)
        if (s[c]) atomicAdd(&counts[c], (unsigned long long)s[c]);
}

size_t stats_smem(int n_c, int d, int l) { return stats_smem_bytes(n_c, d, l); }

int stats_grid(int64_t n, int sm_count) {
    const int64_t tiles = (n + kSThreads - 1) / kSThreads;
    // 3 blocks are resident per SM (166 registers, 62 KB of shared memory): the grid must be a multiple of that,
    // 4 per SM ran 1.33 waves (4.47 ms at 10 M trajectories; 3 / 6 / 12 / 24 per SM: 3.21 / 3.16 / 3.25 / 3.19 ms)
    int per_sm = 6;
    if (const char *e = getenv("MTM_STATS_BLOCKS_PER_SM")) {  // tuning knob
        const int v = atoi(e);
        if (v > 0) per_sm = v;
    }
    const int64_t cap = (int64_t)sm_count * per_sm;
    return (int)(tiles < cap ? (tiles > 0 ? tiles : 1) : cap);
}

template <int D_, int L_>
static cudaError_t launch_stats_dl(const StatsArgs &a, int grid, cudaStream_t s) {
    const size_t smem = stats_smem(a.n_c, a.d, a.l);
    auto k = stats_kernel<D_, L_>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    k<<<dim3(grid, a.R > 1 ? a.R : 1), kSThreads, smem, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_stats(const StatsArgs &a, int grid, cudaStream_t s) {
#define MTM_CASE(DD, LL) \
    if (a.d == DD && a.l == LL) return launch_stats_dl<DD, LL>(a, grid, s);
    MTM_CASE(4, 4)
    MTM_CASE(2, 4)
    MTM_CASE(2, 3)
    MTM_CASE(1, 1)
    MTM_CASE(3, 3)
    MTM_CASE(2, 2)
    MTM_CASE(5, 3)
#undef MTM_CASE
    return launch_stats_dl<0, 0>(a, grid, s);
}

cudaError_t launch_stats_reduce(const double *partials, int grid, int n_vals, double *out, double *scratch,
                                const Ctrl *ctrl, int check_done, int with_switches, double n_local, int R,
                                cudaStream_t s, const StatsCache *cache, double *bases, int grid_delta) {
    const int threads = 128;
    const int S = grid >= 4 * kReduceSlices ? kReduceSlices : 1;
    const int RR = R > 1 ? R : 1;
    const dim3 g1((n_vals + threads - 1) / threads, S, RR);
    stats_reduce1_kernel<<<g1, threads, 0, s>>>(partials, grid, n_vals, scratch, ctrl, check_done,
                                                RR == 1 ? cache : nullptr, grid_delta);
    stats_reduce2_kernel<<<dim3(g1.x, RR), threads, 0, s>>>(scratch, S, n_vals, out, ctrl, check_done, with_switches,
                                                             n_local, RR == 1 ? cache : nullptr, bases);
    return cudaGetLastError();
}

cudaError_t launch_finalize(const FinalizeArgs &a, const double *tail, cudaStream_t s) {
    // (28 us per launch on the critical path of every iteration: serial Cholesky / division chains of 3 n_c threads.
    // Staging the statistics in shared memory first was measured and changed nothing: it is not load latency.)
    const int blocks = a.R > 1 ? a.R : 1;
    int threads = 32;
    while (threads < 3 * a.n_c && threads < 256) threads *= 2;
#define MTM_CASE(DD, LL)                                             \
    if (a.d == DD && a.l == LL) {                                    \
        finalize_kernel<DD, LL><<<blocks, threads, 0, s>>>(a, tail); \
        return cudaGetLastError();                                   \
    }
    MTM_CASE(4, 4)
    MTM_CASE(2, 4)
    MTM_CASE(2, 3)
    MTM_CASE(1, 1)
    MTM_CASE(3, 3)
    MTM_CASE(2, 2)
    MTM_CASE(5, 3)
#undef MTM_CASE
    finalize_kernel<0, 0><<<blocks, threads, 0, s>>>(a, tail);
    return cudaGetLastError();
}

cudaError_t launch_counts(const int32_t *assign, int64_t n, int n_c, unsigned long long *counts, cudaStream_t s) {
    int grid = (int)((n + 255) / 256);
    if (grid > 1184) grid = 1184;
    if (grid < 1) grid = 1;
    counts_kernel<<<grid, 256, 0, s>>>(assign, n, n_c, counts);
    return cudaGetLastError();
}

}  // namespace mtm
