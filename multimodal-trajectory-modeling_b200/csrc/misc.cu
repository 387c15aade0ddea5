// Small device helpers that are not part of the EM arithmetic.
#include "common.cuh"

namespace mtm {

// ---- data-set fingerprint --------------------------------------------------------------------------------------
// 128-bit key of a (T, n, w) float64 array: the wrapping sum over all elements of two mixes of (bit pattern,
// position in the CALLER's order).  Integer sums commute, so the key does not depend on the reduction order, on
// the grid, or on the storage order of the resident copy (perm maps resident position -> caller's index).  It
// stands in for the reference's md5 over the data bytes (MMM:177-188) as the cache key of large data sets:
// one pass at HBM speed instead of ~10 s of host md5 for 5 GB.  Not cryptographic.
__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {  // splitmix64 finaliser
    x ^= x >> 30;
    x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27;
    x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}

__global__ void __launch_bounds__(256) fingerprint_kernel(const double *__restrict__ a, int64_t T, int64_t n, int w,
                                                          const int64_t *__restrict__ perm, unsigned long long seed,
                                                          unsigned long long *out2) {
    unsigned long long h0 = 0, h1 = 0;
    const int64_t rows = T * n;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = r / n, p = r - t * n;
        const int64_t i = perm ? perm[p] : p;
        const unsigned long long base = (unsigned long long)((t * n + i) * w);
        for (int j = 0; j < w; ++j) {
            const unsigned long long bits = (unsigned long long)__double_as_longlong(__ldg(a + r * w + j));
            const unsigned long long pos = mix64(base + j + seed);
            h0 += mix64(bits ^ pos);
            h1 += mix64((bits << 32 | bits >> 32) + pos * 0x9e3779b97f4a7c15ull);
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        h0 += __shfl_xor_sync(0xffffffffu, h0, o);
        h1 += __shfl_xor_sync(0xffffffffu, h1, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(out2, h0);
        atomicAdd(out2 + 1, h1);
    }
}

cudaError_t launch_fingerprint(const double *a, int64_t T, int64_t n, int w, const int64_t *perm,
                               unsigned long long seed, unsigned long long *out2, int sm_count, cudaStream_t s) {
    const int64_t rows = T * n;
    int grid = (int)((rows + 255) / 256);
    if (grid > sm_count * 8) grid = sm_count * 8;
    if (grid < 1) grid = 1;
    fingerprint_kernel<<<grid, 256, 0, s>>>(a, T, n, w, perm, seed, out2);
    return cudaGetLastError();
}

// ---- label hash + statistics-cache decision (common.cuh StatsCache) ------------------------------------------------
// 128-bit order-independent hash of a label vector (position and label mixed, wrapping sums), followed — in the
// block that finishes last — by the cache bookkeeping of one fused EM iteration:
//   role 0, on the labels the E-step starts from: which cached entry holds their statistics, if any, and whether
//           the iteration runs as delta statistics;
//   role 1, on the labels it produced: publish the entry the reduction just wrote.
__global__ void __launch_bounds__(256) label_hash_kernel(const int32_t *__restrict__ assign, int64_t n, int n_c,
                                                         StatsCache *cache, const Ctrl *ctrl, int check_done, int role,
                                                         int allow_delta) {
    if (check_done && ctrl->done) return;
    unsigned long long h0 = 0, h1 = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long key =
            (unsigned long long)i * 0x9e3779b97f4a7c15ull + (unsigned long long)(unsigned int)__ldg(assign + i);
        const unsigned long long m = mix64(key + 0x632be59bd9b4e019ull);
        h0 += m;
        h1 += (m ^ (m >> 29)) * 0xd6e8feb86659fd93ull + key;
    }
    for (int o = 16; o > 0; o >>= 1) {
        h0 += __shfl_xor_sync(0xffffffffu, h0, o);
        h1 += __shfl_xor_sync(0xffffffffu, h1, o);
    }
    // one pair of atomics per BLOCK: thousands of same-address atomics serialise in L2 (19 us at 1.25 M labels with
    // one pair per warp)
    __shared__ unsigned long long s_h[2][8];
    if ((threadIdx.x & 31) == 0) {
        s_h[0][threadIdx.x >> 5] = h0;
        s_h[1][threadIdx.x >> 5] = h1;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    for (int w = 1; w < 8; ++w) {
        h0 += s_h[0][w];
        h1 += s_h[1][w];
    }
    atomicAdd(&cache->acc0, h0);
    atomicAdd(&cache->acc1, h1);
    __threadfence();
    if (atomicAdd(&cache->ticket, 1u) != gridDim.x - 1) return;
    __threadfence();
    const unsigned long long a0 = atomicAdd(&cache->acc0, 0ull), a1 = atomicAdd(&cache->acc1, 0ull);
    if (role == 0) {
        int match = -1;
        for (int k = 0; k < 2; ++k) {
            const StatsCacheEntry &e = cache->e[k];
            if (e.valid && e.n == n && e.n_c == n_c && e.h0 == a0 && e.h1 == a1) match = k;
        }
        int mode = 0, src = 0, dst = 0;
        if (match >= 0) {
            src = match;
            dst = 1 - match;  // the matching entry stays: a caller that steps from these labels again hits it
            if (allow_delta && cache->e[match].sw * 8 <= n) mode = 1;
        } else {
            dst = !cache->e[0].valid ? 0 : (!cache->e[1].valid ? 1 : (cache->e[0].stamp <= cache->e[1].stamp ? 0 : 1));
        }
        cache->e[dst].valid = 0;
        cache->mode = mode;
        cache->src = src;
        cache->dst = dst;
        cache->next_chunk = 0;
    } else {
        StatsCacheEntry &e = cache->e[cache->dst];
        e.h0 = a0;
        e.h1 = a1;
        e.n = n;
        e.n_c = n_c;
        e.sw = (long long)ctrl->n_switches;
        e.stamp = ++cache->clock;
        e.valid = 1;
    }
    cache->acc0 = 0;
    cache->acc1 = 0;
    cache->ticket = 0;
}

cudaError_t launch_label_hash(const int32_t *assign, int64_t n, int n_c, StatsCache *cache, const Ctrl *ctrl,
                              int check_done, int role, int allow_delta, int sm_count, cudaStream_t s) {
    int grid = (int)((n + 2047) / 2048);  // at least eight labels per thread
    if (grid > sm_count * 4) grid = sm_count * 4;
    if (grid < 1) grid = 1;
    label_hash_kernel<<<grid, 256, 0, s>>>(assign, n, n_c, cache, ctrl, check_done, role, allow_delta);
    return cudaGetLastError();
}

}  // namespace mtm
