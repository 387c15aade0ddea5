// Device-side twin of the reference's random initial assignment:
//     np.random.default_rng(seed).integers(low=0, high=n_c, size=n)
// (framework/marginalizable_mixture_model.py:128, 166-168), bit for bit.
//
// numpy's algorithm (numpy/random/_pcg64.pyx, src/pcg64/pcg64.h, src/distributions/distributions.c):
//  * PCG64 = 128-bit LCG  state <- state * M + inc,  output XSL-RR of the NEW state;
//  * integers() with a range that fits 32 bits draws 32-bit numbers through the bit generator's
//    buffered next_uint32: 64-bit output k gives draw 2k (low half) and draw 2k+1 (high half);
//  * each element uses Lemire's multiply-shift: m = draw * n_c, result m >> 32, and REJECTS the draw when
//    (m & 0xffffffff) < (2^32 mod n_c) — which shifts every later element by one draw.
// The host supplies the 128-bit (state, inc) of the seeded generator (numpy computes the SeedSequence
// hash).  Pass 1 flags the rejected draws (about n * n_c / 2^32 of them: a handful in 10^9); pass 2 lets
// every thread jump its own copy of the LCG ahead (O(log k) 128-bit multiply-adds) and emit its elements,
// skipping the flagged draws.
#include "common.cuh"

namespace mtm {

typedef unsigned __int128 u128;

__device__ __forceinline__ u128 make128(unsigned long long hi, unsigned long long lo) { return ((u128)hi << 64) | lo; }

__device__ __forceinline__ u128 pcg_mult() { return make128(0x2360ED051FC65DA4ULL, 0x4385DF649FCCF645ULL); }

// state after `delta` steps (Brown, "Random number generation with arbitrary strides")
__device__ u128 pcg_advance(u128 state, u128 inc, unsigned long long delta) {
    u128 cur_mult = pcg_mult(), cur_plus = inc, acc_mult = 1, acc_plus = 0;
    while (delta > 0) {
        if (delta & 1) {
            acc_mult *= cur_mult;
            acc_plus = acc_plus * cur_mult + cur_plus;
        }
        cur_plus = (cur_mult + 1) * cur_plus;
        cur_mult *= cur_mult;
        delta >>= 1;
    }
    return acc_mult * state + acc_plus;
}

__device__ __forceinline__ unsigned long long pcg_output(u128 s) {
    const unsigned long long hi = (unsigned long long)(s >> 64), lo = (unsigned long long)s;
    const unsigned long long x = hi ^ lo;
    const unsigned int rot = (unsigned int)(hi >> 58);
    return (x >> rot) | (x << ((64 - rot) & 63));
}

constexpr int kDrawsPerThread = 64;  // 32 consecutive 64-bit outputs per thread

// 32-bit draw k of the stream: 64-bit output k/2 (the generator steps before it outputs), low half first
struct DrawStream {
    u128 state, inc;
    unsigned long long cur;
    long long k;  // index of the next 32-bit draw
    __device__ void seek(u128 s0, u128 inc_, long long k0) {
        inc = inc_;
        k = k0;
        state = pcg_advance(s0, inc_, (unsigned long long)(k0 / 2) + 1);
        cur = pcg_output(state);
    }
    __device__ unsigned int next() {
        unsigned int v;
        if ((k & 1) == 0) {
            v = (unsigned int)cur;
        } else {
            v = (unsigned int)(cur >> 32);
            state = state * pcg_mult() + inc;
            cur = pcg_output(state);
        }
        ++k;
        return v;
    }
};

// pass 1: list the indices of the rejected draws among the first n_draws (unordered; tiny)
__global__ void rng_reject_kernel(unsigned long long s_hi, unsigned long long s_lo, unsigned long long i_hi,
                                  unsigned long long i_lo, unsigned int n_c, long long n_draws,
                                  long long *rej_list, int rej_cap, int *rej_count) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long k0 = t * kDrawsPerThread;
    if (k0 >= n_draws) return;
    const unsigned int threshold = (unsigned int)((0x100000000ULL) % n_c);
    if (threshold == 0) return;  // power of two: nothing is ever rejected
    DrawStream ds;
    ds.seek(make128(s_hi, s_lo), make128(i_hi, i_lo), k0);
    for (int j = 0; j < kDrawsPerThread && k0 + j < n_draws; ++j) {
        const unsigned long long m = (unsigned long long)ds.next() * n_c;
        if ((unsigned int)m < threshold) {
            const int slot = atomicAdd(rej_count, 1);
            if (slot < rej_cap) rej_list[slot] = k0 + j;
        }
    }
}

// pass 2: element i = the i-th ACCEPTED draw.  rej_sorted holds the n_rej rejected draw indices ascending.
__global__ void rng_emit_kernel(unsigned long long s_hi, unsigned long long s_lo, unsigned long long i_hi,
                                unsigned long long i_lo, unsigned int n_c, long long n, const long long *rej_sorted,
                                int n_rej, int32_t *out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long e0 = t * kDrawsPerThread;
    if (e0 >= n) return;
    // draw index of element e0: smallest k with (k - #rejected draws below k) == e0 and k itself accepted
    long long k = e0;
    int r = 0;
    while (r < n_rej && rej_sorted[r] <= k) {
        ++r;
        ++k;
    }
    DrawStream ds;
    ds.seek(make128(s_hi, s_lo), make128(i_hi, i_lo), k);
    for (int j = 0; j < kDrawsPerThread && e0 + j < n; ++j) {
        while (r < n_rej && rej_sorted[r] == ds.k) {  // skip rejected draws
            ds.next();
            ++r;
        }
        const unsigned long long m = (unsigned long long)ds.next() * n_c;
        out[e0 + j] = (int32_t)(m >> 32);
    }
}

cudaError_t launch_rng_reject(const unsigned long long st[4], unsigned int n_c, long long n_draws, long long *rej_list,
                              int rej_cap, int *rej_count, cudaStream_t s) {
    const long long threads = (n_draws + kDrawsPerThread - 1) / kDrawsPerThread;
    const int block = 128;
    rng_reject_kernel<<<(unsigned int)((threads + block - 1) / block), block, 0, s>>>(st[0], st[1], st[2], st[3], n_c,
                                                                                       n_draws, rej_list, rej_cap, rej_count);
    return cudaGetLastError();
}

cudaError_t launch_rng_emit(const unsigned long long st[4], unsigned int n_c, long long n, const long long *rej_sorted,
                            int n_rej, int32_t *out, cudaStream_t s) {
    const long long threads = (n + kDrawsPerThread - 1) / kDrawsPerThread;
    const int block = 128;
    rng_emit_kernel<<<(unsigned int)((threads + block - 1) / block), block, 0, s>>>(st[0], st[1], st[2], st[3], n_c, n,
                                                                                     rej_sorted, n_rej, out);
    return cudaGetLastError();
}

}  // namespace mtm
