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
        stats_tile(dm, a.z, a.x, n, a.T, i, valid ? assign[i] : -1, n_c, sm);
    }
    __syncthreads();
    double *out = a.partials + 2 * ((size_t)r * gridDim.x + blockIdx.x) * n_c * sstride;
    for (int k = tid; k < 2 * n_c * sstride; k += kSThreads) out[k] = sm.acc[k];
}

// Two-stage fixed-order reduction of the per-block (hi, lo) partials.
//  stage 1: slice s (blockIdx.y) adds blocks s, s + S, s + 2S, ... in that order  -> scratch[s][v]
//  stage 2: thread v adds the S slices in order                                  -> out[v]
// tail slots: out[n_vals] = n_switches of the preceding E-step (exact integer as double),
//             out[n_vals+1] = local n
__global__ void stats_reduce1_kernel(const double *__restrict__ partials_all, int grid, int n_vals,
                                     double *scratch_all, const Ctrl *ctrl, int check_done) {
    const int r = blockIdx.z;  // batched restarts
    if (check_done && ctrl[r].done) return;
    const double *partials = partials_all + 2 * (size_t)r * grid * n_vals;
    double *scratch = scratch_all + 2 * (size_t)r * gridDim.y * n_vals;
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y, S = gridDim.y;
    if (v >= n_vals) return;
    double hi = 0.0, lo = 0.0;
    for (int b = s; b < grid; b += S) {
        const double2 p = *reinterpret_cast<const double2 *>(partials + 2 * ((size_t)b * n_vals + v));
        dd_add(hi, lo, p.x);
        lo += p.y;
    }
    *reinterpret_cast<double2 *>(scratch + 2 * ((size_t)s * n_vals + v)) = make_double2(hi, lo);
}

__global__ void stats_reduce2_kernel(const double *__restrict__ scratch_all, int S, int n_vals, double *out_all,
                                     const Ctrl *ctrl, int check_done, int with_switches, double n_local) {
    const int r = blockIdx.y;
    if (check_done && ctrl[r].done) return;
    const double *scratch = scratch_all + 2 * (size_t)r * S * n_vals;
    double *out = out_all + (size_t)r * (n_vals + 2);
    const unsigned long long *n_switches_src = with_switches ? &ctrl[r].n_switches : nullptr;
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < n_vals) {
        double hi = 0.0, lo = 0.0;
        for (int s = 0; s < S; ++s) {
            const double2 p = *reinterpret_cast<const double2 *>(scratch + 2 * ((size_t)s * n_vals + v));
            dd_add(hi, lo, p.x);
            lo += p.y;
        }
        out[v] = hi + lo;
    }
    if (v == 0) {
        out[n_vals] = n_switches_src ? (double)(*n_switches_src) : 0.0;
        out[n_vals + 1] = n_local;
    }
}

// ------------------------------------------------------------------------------------------------
// finalise: parameters from moments.  One thread per cluster; run-time dimensions; tiny.
// ------------------------------------------------------------------------------------------------
constexpr int MX = MTM_MAX_DIM;

__device__ inline double qnan() { return __longlong_as_double(0x7ff8000000000000LL); }

// Cholesky solve of (M + alpha I) B = R for SPD M (p x p), R (p x q); M is overwritten by the factor.
// returns false if a pivot is not positive.
__device__ bool chol_solve(double (*M)[MX], double (*R)[MX], int p, int q, double alpha) {
    for (int i = 0; i < p; ++i) M[i][i] += alpha;
    for (int j = 0; j < p; ++j) {
        double s = M[j][j];
        for (int k = 0; k < j; ++k) s -= M[j][k] * M[j][k];
        if (!(s > 0.0)) return false;
        const double dj = sqrt(s);
        M[j][j] = dj;
        for (int i = j + 1; i < p; ++i) {
            double t = M[i][j];
            for (int k = 0; k < j; ++k) t -= M[i][k] * M[j][k];
            M[i][j] = t / dj;
        }
    }
    for (int c = 0; c < q; ++c) {
        for (int i = 0; i < p; ++i) {
            double t = R[i][c];
            for (int k = 0; k < i; ++k) t -= M[i][k] * R[k][c];
            R[i][c] = t / M[i][i];
        }
        for (int i = p - 1; i >= 0; --i) {
            double t = R[i][c];
            for (int k = i + 1; k < p; ++k) t -= M[k][i] * R[k][c];
            R[i][c] = t / M[i][i];
        }
    }
    return true;
}

// moments block [cnt | sum v (K) | upper(sum vv')] of v = [a (p) | b (q)] ->
//   B = argmin |b - a B|^2 + alpha |B|^2 (p x q), Cov = cov(b - a B) with mean removed and N-1
// (regress / regress_alpha, MMM:1167-1233)
__device__ int regress_from_moments(const double *blk, int p, int q, double alpha, double *B_out /*p*q*/,
                                    double *C_out /*q*q*/) {
    const int K = p + q;
    const double cnt = blk[0];
    const double *sv = blk + 1;
    const double *up = blk + 1 + K;
    int status = 0;
    if (!(cnt > 0.0)) {  // empty design: lstsq / Ridge raise in the reference
        for (int k = 0; k < p * q; ++k) B_out[k] = qnan();
        for (int k = 0; k < q * q; ++k) C_out[k] = qnan();
        return MTM_STATUS_EMPTY_FIT;
    }
    double full[2 * MX][2 * MX];
    {
        int o = 0;
        for (int a = 0; a < K; ++a)
            for (int b = a; b < K; ++b) {
                full[a][b] = up[o];
                full[b][a] = up[o];
                ++o;
            }
    }
    double M[MX][MX], B[MX][MX];
    for (int i = 0; i < p; ++i) {
        for (int j = 0; j < p; ++j) M[i][j] = full[i][j];
        for (int c = 0; c < q; ++c) B[i][c] = full[i][p + c];
    }
    if (!chol_solve(M, B, p, q, alpha)) {
        status |= MTM_STATUS_CHOL_FAIL;
        for (int i = 0; i < p; ++i)
            for (int c = 0; c < q; ++c) B[i][c] = qnan();
    }
    // mean residual and residual second moment
    double rbar[MX];
    for (int c = 0; c < q; ++c) {
        double s = sv[p + c];
        for (int i = 0; i < p; ++i) s -= sv[i] * B[i][c];
        rbar[c] = s / cnt;
    }
    // E = Sxy - Sxx B (zero for least squares up to rounding, alpha B for ridge)
    double E[MX][MX];
    for (int i = 0; i < p; ++i)
        for (int c = 0; c < q; ++c) {
            double s = full[i][p + c];
            for (int k = 0; k < p; ++k) s -= full[i][k] * B[k][c];
            E[i][c] = s;
        }
    for (int a = 0; a < q; ++a)
        for (int b = a; b < q; ++b) {
            // SSR = Syy - B'Sxy - E'B
            double s = full[p + a][p + b];
            for (int i = 0; i < p; ++i) s -= B[i][a] * full[i][p + b];
            for (int i = 0; i < p; ++i) s -= E[i][a] * B[i][b];
            double v = (cnt > 1.0) ? (s - cnt * rbar[a] * rbar[b]) / (cnt - 1.0) : qnan();
            C_out[a * q + b] = v;
            C_out[b * q + a] = v;
        }
    for (int i = 0; i < p; ++i)
        for (int c = 0; c < q; ++c) B_out[i * q + c] = B[i][c];
    return status;
}

__global__ void finalize_kernel(const FinalizeArgs a0, const double *tail0) {
    const int r = blockIdx.x;  // batched restarts: one block per fit
    FinalizeArgs a = a0;
    Ctrl *ctrl = a0.ctrl + r;
    if (a.mode == 1 && ctrl->done) return;
    const int d = a.d, l = a.l, n_c = a.n_c;
    const int F1 = 1 + d + d * (d + 1) / 2, F2 = 1 + 2 * d + d * (2 * d + 1);
    const int sstride = stats_stride(d, l), pstride = param_stride(d, l);
    a.stats = a0.stats + (size_t)r * ((size_t)n_c * sstride + (tail0 ? 2 : 0));
    a.params = a0.params + (size_t)r * n_c * pstride;
    a.switch_hist = a0.switch_hist ? a0.switch_hist + (size_t)r * a.n_steps : nullptr;
    const double *tail = tail0 ? a.stats + (size_t)n_c * sstride : nullptr;
    __shared__ int s_go;
    __shared__ int s_status;
    if (threadIdx.x == 0) {
        const double n_total = tail ? tail[1] : (double)a.n_total;
        double mn = 1e300;
        for (int c = 0; c < n_c; ++c) mn = fmin(mn, a.stats[(size_t)c * sstride]);
        (void)n_total;
        int go = 1;
        s_status = 0;
        if (a.mode == 1) {
            const long long sw = (long long)(tail ? tail[0] : 0.0);
            ctrl->min_count = (long long)mn;
            if (ctrl->n_m_steps == 0) {
                // before the first M-step: only the near-empty-cluster guard (MMM:1328-1336)
                if (mn <= 3.0) {
                    ctrl->done = 1;
                    ctrl->exit_code = MTM_EXIT_GUARD0;
                    go = 0;
                }
            } else if (ctrl->status & (MTM_STATUS_NONFINITE | MTM_STATUS_SINGULAR)) {
                // the reference's E-step raises LinAlgError here (det/solve on a non-finite or
                // singular covariance block): stop, the host re-raises
                ctrl->done = 1;
                ctrl->exit_code = MTM_EXIT_ERROR;
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

    for (int c = threadIdx.x; c < n_c; c += blockDim.x) {
        const double *st = a.stats + (size_t)c * sstride;
        double *pr = a.params + (size_t)c * pstride;
        const double n_total = tail ? tail[1] : (double)a.n_total;
        int status = 0;
        pr[0] = st[0] / n_total;  // pi_c = mean(assign == c)
        // initial state: mean and np.cov(ddof=1) of the finite z_1 rows
        const double *b1 = st + 1;
        const double n0 = b1[0];
        double m[MX];
        for (int i = 0; i < d; ++i) {
            m[i] = b1[1 + i] / n0;  // 0/0 = NaN for an empty sample, like np.mean
            pr[1 + i] = m[i];
        }
        {
            int o = 1 + d;
            double *S = pr + 1 + d;
            for (int i = 0; i < d; ++i)
                for (int j = i; j < d; ++j) {
                    const double v = (n0 > 1.0) ? (b1[o] - n0 * m[i] * m[j]) / (n0 - 1.0) : qnan();
                    S[i * d + j] = v;
                    S[j * d + i] = v;
                    ++o;
                }
        }
        double *A = pr + 1 + d + d * d, *G = A + d * d, *H = G + d * d, *Lam = H + d * l;
        status |= regress_from_moments(st + 1 + F1, d, d, a.alpha, A, G);
        status |= regress_from_moments(st + 1 + F1 + F2, d, l, a.alpha, H, Lam);
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
    const int64_t cap = (int64_t)sm_count * 4;
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
                                cudaStream_t s) {
    const int threads = 128;
    const int S = grid >= 4 * kReduceSlices ? kReduceSlices : 1;
    const int RR = R > 1 ? R : 1;
    const dim3 g1((n_vals + threads - 1) / threads, S, RR);
    stats_reduce1_kernel<<<g1, threads, 0, s>>>(partials, grid, n_vals, scratch, ctrl, check_done);
    stats_reduce2_kernel<<<dim3(g1.x, RR), threads, 0, s>>>(scratch, S, n_vals, out, ctrl, check_done, with_switches,
                                                             n_local);
    return cudaGetLastError();
}

cudaError_t launch_finalize(const FinalizeArgs &a, const double *tail, cudaStream_t s) {
    finalize_kernel<<<a.R > 1 ? a.R : 1, 64, 0, s>>>(a, tail);
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
