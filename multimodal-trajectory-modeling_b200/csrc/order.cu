// Pattern-ordered residency: a stable sort of the trajectories by their per-step "how many hidden-state
// coordinates are missing: none / one / more" pattern, and the row permutation kernels that apply it.
//
// Why: when every coordinate of z_t is present the filtered state is a point mass, when one is missing its
// covariance has rank one, and the E-step kernel has shortcuts for both (filter.cuh, "point-mass shortcuts",
// "rank-one shortcuts") — but only on a warp-uniform vote, since a per-lane branch would serialise.  In the
// reference's storage order (framework/marginalizable_mixture_model.py:75-81, trajectories in the caller's order)
// 32 neighbouring trajectories practically never agree at 20 % element-wise missingness; ordered by the pattern
// every aligned group of 32 does (except the groups that straddle a bucket boundary).  The order is a pure function of the data (stable sort, no atomics on the result), so fits
// stay reproducible run to run.
//
// Key of trajectory i: one base-3 digit per time step t < min(T, 8), step 0 the MOST significant:
//   2 = all d coordinates of z[t, i, :] finite (point-mass step), 1 = exactly one missing (rank-one step),
//   0 = anything else.
// Neighbouring buckets therefore differ in their LAST steps first: a warp that straddles a bucket boundary (small
// shards have many small buckets: 3^8 = 6561 of them) still agrees on most of its steps.
// Sort: two stable counting-sort passes over base-81 digits (LSD radix sort: low digit = steps 4..7, then high
// digit = steps 0..3).  Each pass: (1) every warp histograms its contiguous chunk of positions; (2) the
// (bucket-major, warp-minor) table is turned into start offsets (one block per bucket); (3) every warp walks its
// chunk again and hands out positions in order.  perm[p] = original index of the trajectory at resident position p.
#include "common.cuh"

namespace mtm {

constexpr int kOThreads = 256;
constexpr int kOWarps = kOThreads / 32;
constexpr int kBuckets = 256;  // table rows; a base-81 digit uses the first 81
constexpr int kDigit = 81;     // 3^4: four time steps per digit

__device__ __forceinline__ bool finite_hi(double y) { return (__double2hiint(y) & 0x7ff00000) != 0x7ff00000; }

__global__ void __launch_bounds__(256) order_key_kernel(const double *__restrict__ z, int64_t n, int64_t ld, int Tk,
                                                        int d, unsigned short *keys) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int key = 0;
        for (int t = 0; t < Tk; ++t) {
            const double *row = z + ((int64_t)t * ld + i) * d;
            int missing = 0;
            for (int j = 0; j < d; ++j) missing += finite_hi(__ldg(row + j)) ? 0 : 1;
            key = key * 3 + (missing == 0 ? 2 : (missing == 1 ? 1 : 0));
        }
        keys[i] = (unsigned short)key;
    }
}

__device__ __forceinline__ int key_digit(const unsigned short *__restrict__ keys, const int64_t *__restrict__ perm_in,
                                         int64_t pos, int hi, int64_t *idx) {
    const int64_t i = perm_in ? perm_in[pos] : pos;
    *idx = i;
    const int k = keys[i];
    return hi ? k / kDigit : k % kDigit;
}

// chunk of warp `unit`: positions [unit * per, min(n, (unit + 1) * per)), per a multiple of 32
__global__ void __launch_bounds__(kOThreads) order_hist_kernel(const unsigned short *__restrict__ keys,
                                                                const int64_t *__restrict__ perm_in, int hi, int64_t n,
                                                                int64_t per, unsigned int *hist, int units) {
    __shared__ unsigned int s_cnt[kOWarps][kBuckets];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int unit = blockIdx.x * kOWarps + warp;
    for (int b = lane; b < kBuckets; b += 32) s_cnt[warp][b] = 0;
    __syncwarp();
    const int64_t lo = (int64_t)unit * per;
    const int64_t hi_pos = lo + per < n ? lo + per : n;
    for (int64_t base = lo; base < hi_pos; base += 32) {
        const int64_t pos = base + lane;
        const bool valid = pos < hi_pos;
        int64_t idx;
        const int key = valid ? key_digit(keys, perm_in, pos, hi, &idx) : kBuckets;
        const unsigned int m = __match_any_sync(0xffffffffu, key);
        if (valid && (m & ((1u << lane) - 1u)) == 0) s_cnt[warp][key] += __popc(m);
        __syncwarp();
    }
    for (int b = lane; b < kBuckets; b += 32) hist[(size_t)b * units + unit] = s_cnt[warp][b];
}

// Start offsets from the (bucket-major, warp-minor) histogram, in two levels: block b turns row b into its
// exclusive prefix sums (coalesced, block-wide scan per 1024-entry chunk) and records the row total; the scatter
// kernel adds the exclusive sum of the totals of the earlier rows.
__global__ void __launch_bounds__(1024) order_scan_rows_kernel(unsigned int *hist, int units, unsigned int *row_total) {
    __shared__ unsigned int s[1024];
    __shared__ unsigned int s_carry;
    unsigned int *row = hist + (size_t)blockIdx.x * units;
    const int tid = threadIdx.x;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < units; base += 1024) {
        const int k = base + tid;
        const unsigned int v = k < units ? row[k] : 0u;
        s[tid] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {  // Hillis-Steele inclusive scan
            const unsigned int add = tid >= o ? s[tid - o] : 0u;
            __syncthreads();
            s[tid] += add;
            __syncthreads();
        }
        const unsigned int carry = s_carry;
        if (k < units) row[k] = carry + s[tid] - v;
        __syncthreads();
        if (tid == 1023) s_carry = carry + s[1023];
        __syncthreads();
    }
    if (tid == 0) row_total[blockIdx.x] = s_carry;
}

__global__ void __launch_bounds__(kOThreads) order_scatter_kernel(const unsigned short *__restrict__ keys,
                                                                   const int64_t *__restrict__ perm_in, int hi, int64_t n,
                                                                   int64_t per, const unsigned int *__restrict__ offs,
                                                                   const unsigned int *__restrict__ row_total, int units,
                                                                   int64_t *perm_out) {
    __shared__ unsigned int s_pos[kOWarps][kBuckets];
    __shared__ unsigned int s_base[kBuckets];
    static_assert(kOThreads == kBuckets, "one thread per bucket for the base scan");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int unit = blockIdx.x * kOWarps + warp;
    s_base[threadIdx.x] = row_total[threadIdx.x];
    __syncthreads();
    for (int o = 1; o < kBuckets; o <<= 1) {  // inclusive scan of the bucket sizes
        const unsigned int add = threadIdx.x >= o ? s_base[threadIdx.x - o] : 0u;
        __syncthreads();
        s_base[threadIdx.x] += add;
        __syncthreads();
    }
    for (int b = lane; b < kBuckets; b += 32)
        s_pos[warp][b] = offs[(size_t)b * units + unit] + (b ? s_base[b - 1] : 0u);
    __syncwarp();
    const int64_t lo = (int64_t)unit * per;
    const int64_t hi_pos = lo + per < n ? lo + per : n;
    for (int64_t base = lo; base < hi_pos; base += 32) {
        const int64_t pos = base + lane;
        const bool valid = pos < hi_pos;
        int64_t idx = 0;
        const int key = valid ? key_digit(keys, perm_in, pos, hi, &idx) : kBuckets;
        const unsigned int m = __match_any_sync(0xffffffffu, key);
        const int rank = __popc(m & ((1u << lane) - 1u));
        unsigned int start = 0;
        if (valid) start = s_pos[warp][key];
        __syncwarp();
        if (valid && rank == 0) s_pos[warp][key] = start + __popc(m);
        __syncwarp();
        if (valid) perm_out[start + rank] = idx;
    }
}

// (rows, n, w) arrays: gather dst[r, p, :] = src[r, perm[p], :]  or scatter dst[r, perm[p], :] = src[r, p, :]
template <class T>
__global__ void __launch_bounds__(256) permute_rows_kernel(const T *__restrict__ src, T *__restrict__ dst, int64_t rows,
                                                            int64_t n, int w, const int64_t *__restrict__ perm,
                                                            int inverse) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int64_t q = perm[p];
    const int64_t from = inverse ? p : q, to = inverse ? q : p;
    for (int64_t r = 0; r < rows; ++r) {
        const T *s = src + (r * n + from) * w;
        T *o = dst + (r * n + to) * w;
        if (sizeof(T) == 8 && (w & 1) == 0) {  // 16-byte rows: vector copies (rows are 16-byte aligned when w is even)
            for (int k = 0; k < w; k += 2)
                *reinterpret_cast<double2 *>(o + k) = __ldg(reinterpret_cast<const double2 *>(s + k));
        } else {
            for (int k = 0; k < w; ++k) o[k] = __ldg(s + k);
        }
    }
}

int order_units(int sm_count) { return sm_count * 4 * kOWarps; }  // warps of the sort grid

static size_t align256(size_t b) { return (b + 255) & ~size_t(255); }

// [hist | row totals | keys (n x u16) | intermediate permutation (n x i64)]
size_t order_scratch_bytes(int64_t n, int sm_count) {
    return align256((size_t)kBuckets * (order_units(sm_count) + 1) * sizeof(unsigned int)) + align256((size_t)n * 2) +
           (size_t)n * 8 + 256;
}

cudaError_t launch_pattern_order(const double *z, int64_t n, int64_t ld, int T, int d, int sm_count, void *scratch,
                                 int64_t *perm, cudaStream_t s) {
    const int units = order_units(sm_count);
    const int grid = units / kOWarps;
    unsigned char *base = static_cast<unsigned char *>(scratch);
    unsigned int *hist = reinterpret_cast<unsigned int *>(base);
    unsigned int *row_total = hist + (size_t)kBuckets * units;
    base += align256((size_t)kBuckets * (units + 1) * sizeof(unsigned int));
    unsigned short *keys = reinterpret_cast<unsigned short *>(base);
    base += align256((size_t)n * 2);
    int64_t *perm_a = reinterpret_cast<int64_t *>(base);
    int64_t per = (n + units - 1) / units;
    per = (per + 31) / 32 * 32;
    const int Tk = T < 8 ? T : 8;
    int kgrid = (int)((n + 255) / 256);
    if (kgrid > sm_count * 16) kgrid = sm_count * 16;
    order_key_kernel<<<kgrid, 256, 0, s>>>(z, n, ld, Tk, d, keys);
    const bool two = Tk > 4;  // keys below 81 have a zero high digit: one pass sorts them
    int64_t *out_a = two ? perm_a : perm;
    order_hist_kernel<<<grid, kOThreads, 0, s>>>(keys, nullptr, 0, n, per, hist, units);
    order_scan_rows_kernel<<<kBuckets, 1024, 0, s>>>(hist, units, row_total);
    order_scatter_kernel<<<grid, kOThreads, 0, s>>>(keys, nullptr, 0, n, per, hist, row_total, units, out_a);
    if (two) {
        order_hist_kernel<<<grid, kOThreads, 0, s>>>(keys, perm_a, 1, n, per, hist, units);
        order_scan_rows_kernel<<<kBuckets, 1024, 0, s>>>(hist, units, row_total);
        order_scatter_kernel<<<grid, kOThreads, 0, s>>>(keys, perm_a, 1, n, per, hist, row_total, units, perm);
    }
    return cudaGetLastError();
}

cudaError_t launch_permute_f64(const double *src, double *dst, int64_t rows, int64_t n, int w, const int64_t *perm,
                               int inverse, cudaStream_t s) {
    permute_rows_kernel<double><<<(unsigned int)((n + 255) / 256), 256, 0, s>>>(src, dst, rows, n, w, perm, inverse);
    return cudaGetLastError();
}

cudaError_t launch_permute_i32(const int32_t *src, int32_t *dst, int64_t rows, int64_t n, const int64_t *perm,
                               int inverse, cudaStream_t s) {
    permute_rows_kernel<int32_t><<<(unsigned int)((n + 255) / 256), 256, 0, s>>>(src, dst, rows, n, 1, perm, inverse);
    return cudaGetLastError();
}

}  // namespace mtm
