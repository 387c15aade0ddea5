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

}  // namespace mtm
