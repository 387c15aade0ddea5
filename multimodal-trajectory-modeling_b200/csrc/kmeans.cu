// k-means initial assignment on the device.
//
// The reference's constructor (framework/marginalizable_mixture_model.py:132-164) clusters the first non-null state
// of every trajectory (NaN -> column mean) with sklearn.cluster.KMeans(init="k-means++", random_state=seed).  At 10 M
// trajectories x 352 restarts of a model-selection sweep that host call is the wall-clock bottleneck.  These
// kernels follow sklearn's algorithm step for step so that the device start lands where sklearn's does:
//   * the features are centred by their column means (KMeans.fit), tol = mean(var) * 1e-4 (_tolerance);
//   * k-means++ (_kmeans_plusplus): the first centre and the 2 + log(k) uniform draws per further centre come
//     from the caller's numpy RandomState(seed) — the same stream sklearn consumes; a candidate is the first index
//     whose running sum of closest squared distances reaches u * potential (searchsorted on the cumulative sum),
//     the candidate that lowers the potential most is kept;
//   * Lloyd (_kmeans_single_lloyd): labels from the current centres, new centres = member means, stop when no
//     label changed (strict) or the squared centre shift is <= tol, then — unless strictly converged — one more
//     labelling with the final centres.
// Differences are confined to rounding: distances are (x - c)^2 sums here and |c|^2 - 2 x.c in sklearn's GEMM form,
// the cumulative sum is chunked here and sequential in numpy; they matter only for exact ties.
// All reductions have a fixed order (no floating-point atomics): the same labels every run.
#include <cmath>
#include <cstring>

#include "common.cuh"

namespace mtm {

constexpr int kKmChunks = 1024;  // chunks of the trajectory axis for the cumulative sum
constexpr int kKmMaxTrials = 8;  // 2 + log(64) = 6

// first time step at which all hidden-state coordinates are finite (0 if none: np.argmax of all-False), MMM:132-137
__global__ void __launch_bounds__(256) first_rows_kernel(const double *__restrict__ z, int64_t n, int64_t ld, int T, int d,
                                                         double *__restrict__ rows) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int pick = 0;
        for (int t = 0; t < T; ++t) {
            const double *r = z + ((int64_t)t * ld + i) * d;
            bool all = true;
            for (int j = 0; j < d; ++j) all = all && ((__double2hiint(__ldg(r + j)) & 0x7ff00000) != 0x7ff00000);
            if (all) {
                pick = t;
                break;
            }
        }
        const double *r = z + ((int64_t)pick * ld + i) * d;
        for (int j = 0; j < d; ++j) rows[i * d + j] = __ldg(r + j);
    }
}

// column [sum | sum of squares | count] over the FINITE entries: block partials, then a fixed-order reduction
__global__ void __launch_bounds__(256) col_moments_kernel(const double *__restrict__ x, int64_t n, int w, double *partials) {
    extern __shared__ double s_cm[];  // [8 warps][3 w]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int j = 0; j < w; ++j) {
        double s = 0.0, q = 0.0, c = 0.0;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
            const double v = x[i * w + j];
            const bool f = (__double2hiint(v) & 0x7ff00000) != 0x7ff00000;
            s += f ? v : 0.0;
            q += f ? v * v : 0.0;
            c += f ? 1.0 : 0.0;
        }
        for (int o = 16; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            q += __shfl_xor_sync(0xffffffffu, q, o);
            c += __shfl_xor_sync(0xffffffffu, c, o);
        }
        if (lane == 0) {
            s_cm[(warp * 3 + 0) * w + j] = s;
            s_cm[(warp * 3 + 1) * w + j] = q;
            s_cm[(warp * 3 + 2) * w + j] = c;
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 3 * w; k += blockDim.x) {
        double t = 0.0;
        for (int wp = 0; wp < 8; ++wp) t += s_cm[wp * 3 * w + k];
        partials[(size_t)blockIdx.x * 3 * w + k] = t;
    }
}

__global__ void sum_partials_kernel(const double *__restrict__ partials, int grid, int n_vals, double *out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_vals) return;
    double t = 0.0;
    for (int b = 0; b < grid; ++b) t += partials[(size_t)b * n_vals + k];
    out[k] = t;
}

// x <- (finite ? x : fill) - shift   per column (imputation with the column mean, then centring)
__global__ void __launch_bounds__(256) impute_shift_kernel(double *__restrict__ x, int64_t n, int w, const double *fill,
                                                           const double *shift) {
    const int64_t total = n * w;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(e % w);
        double v = x[e];
        if ((__double2hiint(v) & 0x7ff00000) == 0x7ff00000) v = fill[j];
        x[e] = v - shift[j];
    }
}

__device__ __forceinline__ double sqdist(const double *__restrict__ a, const double *__restrict__ b, int w) {
    double s = 0.0;
    for (int j = 0; j < w; ++j) {
        const double t = a[j] - b[j];
        s = fma(t, t, s);
    }
    return s;
}

// closest[i] = |x_i - x_first|^2
__global__ void __launch_bounds__(256) kpp_first_kernel(const double *__restrict__ x, int64_t n, int w, int64_t first,
                                                        double *__restrict__ closest, double *centers) {
    __shared__ double c[64];
    for (int j = threadIdx.x; j < w; j += blockDim.x) c[j] = x[first * w + j];
    __syncthreads();
    if (blockIdx.x == 0)
        for (int j = threadIdx.x; j < w; j += blockDim.x) centers[j] = c[j];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        closest[i] = sqdist(x + i * w, c, w);
}

// sums of `closest` over kKmChunks contiguous chunks (one block per chunk, fixed order inside)
__global__ void __launch_bounds__(256) chunk_sums_kernel(const double *__restrict__ v, int64_t n, int64_t per, double *sums) {
    __shared__ double s[256];
    const int64_t lo = (int64_t)blockIdx.x * per, hi = lo + per < n ? lo + per : n;
    double t = 0.0;
    for (int64_t i = lo + threadIdx.x; i < hi; i += 256) t += v[i];
    s[threadIdx.x] = t;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) sums[blockIdx.x] = s[0];
}

// candidate l = first index whose running sum of `closest` reaches u_l * potential (np.searchsorted on np.cumsum,
// clipped to n - 1); thread l walks the chunk sums, then its chunk
__global__ void kpp_pick_kernel(const double *__restrict__ closest, int64_t n, int64_t per, const double *__restrict__ sums,
                                int n_chunks, const double *__restrict__ u, int L, double *pot_out, long long *cand) {
    const int l = threadIdx.x;
    double pot = 0.0;
    for (int k = 0; k < n_chunks; ++k) pot += sums[k];
    if (l == 0) *pot_out = pot;
    if (l >= L) return;
    const double target = u[l] * pot;
    double acc = 0.0;
    int k = 0;
    for (; k < n_chunks - 1; ++k) {
        if (acc + sums[k] >= target) break;
        acc += sums[k];
    }
    const int64_t lo = (int64_t)k * per, hi = lo + per < n ? lo + per : n;
    int64_t idx = hi - 1;
    for (int64_t i = lo; i < hi; ++i) {
        acc += closest[i];
        if (acc >= target) {
            idx = i;
            break;
        }
    }
    if (idx > n - 1) idx = n - 1;
    if (idx < 0) idx = 0;
    cand[l] = idx;
}

// potential of every candidate: sum_i min(closest_i, |x_i - x_cand|^2)   -> block partials (L per block)
__global__ void __launch_bounds__(256) kpp_eval_kernel(const double *__restrict__ x, int64_t n, int w,
                                                       const double *__restrict__ closest, const long long *__restrict__ cand,
                                                       int L, double *partials) {
    __shared__ double c[kKmMaxTrials][64];
    __shared__ double red[kKmMaxTrials][8];
    for (int k = threadIdx.x; k < L * w; k += blockDim.x) c[k / w][k % w] = x[cand[k / w] * w + (k % w)];
    __syncthreads();
    double pot[kKmMaxTrials];
#pragma unroll
    for (int l = 0; l < kKmMaxTrials; ++l) pot[l] = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double cl = closest[i];
#pragma unroll
        for (int l = 0; l < kKmMaxTrials; ++l)
            if (l < L) pot[l] += fmin(cl, sqdist(x + i * w, c[l], w));
    }
#pragma unroll
    for (int l = 0; l < kKmMaxTrials; ++l) {
        for (int o = 16; o > 0; o >>= 1) pot[l] += __shfl_xor_sync(0xffffffffu, pot[l], o);
        if ((threadIdx.x & 31) == 0) red[l][threadIdx.x >> 5] = pot[l];
    }
    __syncthreads();
    if (threadIdx.x < L) {
        double t = 0.0;
        for (int wp = 0; wp < 8; ++wp) t += red[threadIdx.x][wp];
        partials[(size_t)blockIdx.x * L + threadIdx.x] = t;
    }
}

// keep candidate `best`: closest <- min(closest, dist), centre c <- x_best
__global__ void __launch_bounds__(256) kpp_commit_kernel(const double *__restrict__ x, int64_t n, int w,
                                                         double *__restrict__ closest, const long long *__restrict__ cand,
                                                         int best, double *center_out) {
    __shared__ double c[64];
    const long long b = cand[best];
    for (int j = threadIdx.x; j < w; j += blockDim.x) c[j] = x[b * w + j];
    __syncthreads();
    if (blockIdx.x == 0)
        for (int j = threadIdx.x; j < w; j += blockDim.x) center_out[j] = c[j];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        closest[i] = fmin(closest[i], sqdist(x + i * w, c, w));
}

// labels from the centres (first minimum wins, like np.argmin); counts the labels that changed
__global__ void __launch_bounds__(256) km_assign_kernel(const double *__restrict__ x, int64_t n, int w,
                                                        const double *__restrict__ centers, int k, int32_t *__restrict__ labels,
                                                        unsigned long long *changed) {
    extern __shared__ double s_c[];  // k * w
    for (int e = threadIdx.x; e < k * w; e += blockDim.x) s_c[e] = centers[e];
    __syncthreads();
    unsigned int ch = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double best = sqdist(x + i * w, s_c, w);
        int arg = 0;
        for (int c = 1; c < k; ++c) {
            const double dd = sqdist(x + i * w, s_c + c * w, w);
            if (dd < best) {
                best = dd;
                arg = c;
            }
        }
        ch += (labels[i] != arg) ? 1u : 0u;
        labels[i] = arg;
    }
    for (int o = 16; o > 0; o >>= 1) ch += __shfl_xor_sync(0xffffffffu, ch, o);
    if ((threadIdx.x & 31) == 0 && ch) atomicAdd(changed, (unsigned long long)ch);
}

// new centres = member means (an empty cluster keeps its centre), squared shift of all centres -> out[0]
__global__ void km_update_kernel(const double *__restrict__ sums_counts, int k, int w, double *centers, double *shift_out,
                                 int *empty_out) {
    __shared__ double sh[64];
    double part = 0.0;
    int empty = 0;
    for (int c = threadIdx.x; c < k; c += blockDim.x) {
        double s2 = 0.0;
        for (int j = 0; j < w; ++j) {
            const double cnt = sums_counts[(size_t)c * 2 * w + w + j];
            const double old = centers[c * w + j];
            const double nw = cnt > 0.0 ? sums_counts[(size_t)c * 2 * w + j] / cnt : old;
            empty |= cnt > 0.0 ? 0 : 1;
            centers[c * w + j] = nw;
            s2 = fma(nw - old, nw - old, s2);
        }
        const double nrm = sqrt(s2);  // sklearn keeps the per-centre norm and squares it again
        part += nrm * nrm;
    }
    sh[threadIdx.x] = part;
    if (empty) atomicOr(empty_out, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int c = 0; c < blockDim.x; ++c) t += sh[c];
        *shift_out = t;
    }
}

static int km_grid(int64_t n, int sm_count) {
    const int64_t b = (n + 255) / 256, cap = (int64_t)sm_count * 8;
    return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

cudaError_t launch_first_rows(const double *z, int64_t n, int64_t ld, int T, int d, double *rows, int sm_count,
                              cudaStream_t s) {
    first_rows_kernel<<<km_grid(n, sm_count), 256, 0, s>>>(z, n, ld, T, d, rows);
    return cudaGetLastError();
}

// scratch layout (doubles): [closest n | chunk sums 1024 | partials grid*max(3w, L) | small 4*64 + 16 | cand 8]
size_t kmeans_scratch_doubles(int64_t n, int w, int n_c, int sm_count) {
    const size_t grid = (size_t)sm_count * 8;
    const size_t part = grid * (size_t)(3 * w > kKmMaxTrials ? 3 * w : kKmMaxTrials);
    return (size_t)n + kKmChunks + part + 4 * 64 + 32 + nanmean_scratch_doubles(n_c, w, sm_count) + (size_t)n_c * 2 * w;
}

// The whole fit.  rows (n, w) is overwritten by its imputed, centred copy.  Synchronises (host decisions).
cudaError_t run_kmeans(double *rows, int64_t n, int w, int n_c, int64_t first_center, const double *u_host, int n_trials,
                       int max_iter, double rel_tol, int32_t *labels, double *centers /* (n_c, w) device */,
                       double *scratch, int sm_count, cudaStream_t s, int *n_iter_out, int *how_out, int *empty_out,
                       double *tol_out, int64_t *launches) {
    const int grid = km_grid(n, sm_count);
    double *closest = scratch;
    double *chunk_sums = closest + n;
    double *partials = chunk_sums + kKmChunks;
    double *small = partials + (size_t)sm_count * 8 * (size_t)(3 * w > kKmMaxTrials ? 3 * w : kKmMaxTrials);
    double *d_mean = small, *d_fill = small + 64, *d_u = small + 128, *d_pots = small + 192, *d_misc = small + 256;
    long long *d_cand = reinterpret_cast<long long *>(d_misc + 8);
    unsigned long long *d_changed = reinterpret_cast<unsigned long long *>(d_misc + 2);
    int *d_empty = reinterpret_cast<int *>(d_misc + 4);
    double *nm_out = small + 4 * 64 + 32;
    double *nm_scratch = nm_out + (size_t)n_c * 2 * w;
    cudaError_t e;
#define KM(expr)                         \
    do {                                 \
        e = (expr);                      \
        if (e != cudaSuccess) return e;  \
    } while (0)
    double host[3 * 64];
    // 1. imputation with the column means of the finite entries (MMM:138-146), centring, tolerance
    col_moments_kernel<<<grid, 256, (size_t)8 * 3 * w * sizeof(double), s>>>(rows, n, w, partials);
    sum_partials_kernel<<<(3 * w + 127) / 128, 128, 0, s>>>(partials, grid, 3 * w, small);
    *launches += 2;
    // (the reduced moments land at `small`; read them back, derive fill / mean / tol on the host)
    KM(cudaMemcpyAsync(host, small, sizeof(double) * 3 * w, cudaMemcpyDeviceToHost, s));
    KM(cudaStreamSynchronize(s));
    double fill[64], mean[64], tol = 0.0;
    for (int j = 0; j < w; ++j) {
        const double sm = host[j], sq = host[w + j], cnt = host[2 * w + j];
        fill[j] = cnt > 0.0 ? sm / cnt : 0.0;
        // after imputation every row has the column mean where it was missing: the mean over all n rows is unchanged
        mean[j] = fill[j];
        const double var_fin = cnt > 0.0 ? sq / cnt - fill[j] * fill[j] : 0.0;  // variance of the finite entries
        tol += (var_fin > 0.0 ? var_fin : 0.0) * (cnt / (double)n);             // imputed entries sit on the mean
    }
    tol = tol / w * rel_tol;
    KM(cudaMemcpyAsync(d_fill, fill, sizeof(double) * w, cudaMemcpyHostToDevice, s));
    KM(cudaMemcpyAsync(d_mean, mean, sizeof(double) * w, cudaMemcpyHostToDevice, s));
    impute_shift_kernel<<<grid, 256, 0, s>>>(rows, n, w, d_fill, d_mean);
    *launches += 1;
    // 2. k-means++
    const int L = n_trials < kKmMaxTrials ? n_trials : kKmMaxTrials;
    const int64_t per = (n + kKmChunks - 1) / kKmChunks;
    kpp_first_kernel<<<grid, 256, 0, s>>>(rows, n, w, first_center, closest, centers);
    *launches += 1;
    for (int c = 1; c < n_c; ++c) {
        KM(cudaMemcpyAsync(d_u, u_host + (size_t)(c - 1) * n_trials, sizeof(double) * L, cudaMemcpyHostToDevice, s));
        chunk_sums_kernel<<<kKmChunks, 256, 0, s>>>(closest, n, per, chunk_sums);
        kpp_pick_kernel<<<1, 32, 0, s>>>(closest, n, per, chunk_sums, kKmChunks, d_u, L, d_misc, d_cand);
        kpp_eval_kernel<<<grid, 256, 0, s>>>(rows, n, w, closest, d_cand, L, partials);
        sum_partials_kernel<<<1, 128, 0, s>>>(partials, grid, L, d_pots);
        *launches += 4;
        double pots[kKmMaxTrials];
        KM(cudaMemcpyAsync(pots, d_pots, sizeof(double) * L, cudaMemcpyDeviceToHost, s));
        KM(cudaStreamSynchronize(s));
        int best = 0;
        for (int l = 1; l < L; ++l)
            if (pots[l] < pots[best]) best = l;  // np.argmin: first minimum
        kpp_commit_kernel<<<grid, 256, 0, s>>>(rows, n, w, closest, d_cand, best, centers + (size_t)c * w);
        *launches += 1;
    }
    // 3. Lloyd
    KM(cudaMemsetAsync(labels, 0xff, sizeof(int32_t) * n, s));  // labels_old = -1
    KM(cudaMemsetAsync(d_empty, 0, sizeof(int), s));
    int it = 0, how = 0;
    const size_t smem_c = (size_t)n_c * w * sizeof(double);
    if (smem_c > 48 * 1024) KM(cudaFuncSetAttribute(km_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
    for (; it < max_iter; ++it) {
        KM(cudaMemsetAsync(d_changed, 0, sizeof(unsigned long long), s));
        km_assign_kernel<<<grid, 256, smem_c, s>>>(rows, n, w, centers, n_c, labels, d_changed);
        KM(launch_nanmean(rows, n, w, labels, n_c, sm_count, nm_scratch, nm_out, nullptr, 0, s));
        km_update_kernel<<<1, 64, 0, s>>>(nm_out, n_c, w, centers, d_misc, d_empty);
        *launches += 4;
        double shift;
        unsigned long long changed;
        KM(cudaMemcpyAsync(&shift, d_misc, sizeof(double), cudaMemcpyDeviceToHost, s));
        KM(cudaMemcpyAsync(&changed, d_changed, sizeof(changed), cudaMemcpyDeviceToHost, s));
        KM(cudaStreamSynchronize(s));
        if (changed == 0ull) {
            how = 1;  // strict convergence: the labels already belong to these centres
            ++it;
            break;
        }
        if (shift <= tol) {
            how = 2;
            ++it;
            break;
        }
    }
    if (how != 1) {  // labels of the final centres
        KM(cudaMemsetAsync(d_changed, 0, sizeof(unsigned long long), s));
        km_assign_kernel<<<grid, 256, smem_c, s>>>(rows, n, w, centers, n_c, labels, d_changed);
        *launches += 1;
    }
    int empty = 0;
    KM(cudaMemcpyAsync(&empty, d_empty, sizeof(int), cudaMemcpyDeviceToHost, s));
    KM(cudaStreamSynchronize(s));
#undef KM
    *n_iter_out = it;
    *how_out = how;
    *empty_out = empty;
    *tol_out = tol;
    return cudaGetLastError();
}

}  // namespace mtm
