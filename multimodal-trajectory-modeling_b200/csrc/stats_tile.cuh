// Sufficient statistics of one 128-trajectory tile, reduced by cluster into the block's shared-memory
// accumulators.  Shared by the stand-alone statistics kernel (mstep.cu) and the fused E-step
// (estep.cu): see the header comment of mstep.cu for the method.
#pragma once
#include "common.cuh"
#include "filter.cuh"

namespace mtm {

// (hi, lo) += x, error-free (Knuth two-sum)
__device__ __forceinline__ void dd_add(double &hi, double &lo, double x) {
    const double t = hi + x;
    const double bp = t - hi;
    const double e = (hi - (t - bp)) + (x - bp);
    hi = t;
    lo += e;
}

template <int K>
struct Moments {  // [count | sum v (K) | sum v v' upper (K(K+1)/2)]
    static constexpr int F = 1 + K + K * (K + 1) / 2;
};

// acc += moments of v (all K components finite) else nothing
template <int KM>
__device__ __forceinline__ void add_moments(double *acc, const double (&v)[KM], int K, bool ok) {
    acc[0] += ok ? 1.0 : 0.0;
    double w[KM];
    MTM_UNROLL
    for (int a = 0; a < KM; ++a)
        if (a < K) w[a] = ok ? v[a] : 0.0;
    MTM_UNROLL
    for (int a = 0; a < KM; ++a)
        if (a < K) acc[1 + a] += w[a];
    int o = 1 + K;
    MTM_UNROLL
    for (int a = 0; a < KM; ++a) {
        if (a < K) {
            MTM_UNROLL
            for (int b = a; b < KM; ++b)
                if (b < K) acc[o + (b - a)] = fma(w[a], w[b], acc[o + (b - a)]);
            o += K - a;
        }
    }
}

template <int NM>
__device__ __forceinline__ bool all_finite(const double (&v)[NM], int n) {
    bool ok = true;
    MTM_UNROLL
    for (int a = 0; a < NM; ++a)
        if (a < n) ok = ok && finite_bits(v[a]);
    return ok;
}

// block-level reduce-by-cluster of per-trajectory features that sit in registers
//   feat[f] (f < F) of thread tid -> s_feat[f * ST + tid]; task (c, f) sums the members of c.
// block-level reduce-by-cluster of per-trajectory features that sit in registers.
//   `pos` = the thread's position in the tile's cluster-sorted order (-1: not counted).  The features
//   are staged ALREADY SORTED (s_feat[f * ST + pos]), so the members of cluster c are the contiguous
//   range [start[c], start[c+1]) of every feature row: task (c, f) adds them with independent,
//   address-predictable loads (no index indirection), four interleaved chains in a fixed pattern.
template <int FM>
__device__ __forceinline__ void reduce_by_cluster(const double *feat, int F, double *s_feat, int pos,
                                                  const int *s_start, int n_c, double *s_acc, int sstride,
                                                  int gofs) {
    constexpr int ST = kSThreads + 1;
    const int tid = threadIdx.x;
    if (pos >= 0) {
        MTM_UNROLL
        for (int f = 0; f < FM; ++f)
            if (f < F) s_feat[f * ST + pos] = feat[f];
    }
    __syncthreads();
    for (int task = tid; task < n_c * F; task += kSThreads) {
        const int c = task / F, f = task - c * F;
        const int b = s_start[c], e = s_start[c + 1];
        if (e > b) {
            const double *row = s_feat + f * ST;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            int r = b;
            for (; r + 3 < e; r += 4) {
                s0 += row[r];
                s1 += row[r + 1];
                s2 += row[r + 2];
                s3 += row[r + 3];
            }
            for (; r < e; ++r) s0 += row[r];
            double *acc = s_acc + 2 * ((size_t)c * sstride + gofs + f);
            dd_add(acc[0], acc[1], (s0 + s1) + (s2 + s3));
        }
    }
    __syncthreads();
}

// shared-memory carve-up of the statistics part of a block
struct StatsSmem {
    double *acc;   // n_c * sstride * 2 (hi, lo)
    double *feat;  // FMAX * (kSThreads + 1)
    int *start;    // n_c + 1
    int *wcnt;     // warps * n_c
    __device__ __forceinline__ void carve(unsigned char *base, int n_c, int d, int l) {
        const int F2 = 1 + 2 * d + d * (2 * d + 1), F3 = 1 + (d + l) + (d + l) * (d + l + 1) / 2;
        const int FMAX = F2 > F3 ? F2 : F3;
        acc = reinterpret_cast<double *>(base);
        feat = acc + 2 * (size_t)n_c * stats_stride(d, l);
        start = reinterpret_cast<int *>(feat + (size_t)FMAX * (kSThreads + 1));
        wcnt = start + (n_c + 1);
    }
};

__host__ __device__ inline size_t stats_smem_bytes(int n_c, int d, int l) {
    const int F2 = 1 + 2 * d + d * (2 * d + 1), F3 = 1 + (d + l) + (d + l) * (d + l + 1) / 2;
    const int FMAX = F2 > F3 ? F2 : F3;
    return (size_t)2 * n_c * stats_stride(d, l) * 8 + (size_t)FMAX * (kSThreads + 1) * 8 +
           (size_t)(n_c + 1) * 4 + (size_t)(kSThreads / 32) * n_c * 4 + 16;
}

// z_t - pivot: the moments are accumulated about a common point near the data's mean (the per-feature mean of
// the finite entries, api.cu ensure_pivot), which removes the cancellation of one-pass raw moments when
// |mean| >> spread; finalize_kernel works with the centred moments directly.  NaN / inf stay non-finite.
template <int NM>
__device__ __forceinline__ void centre_row(double (&v)[NM], int n, const double *piv) {
    MTM_UNROLL
    for (int k = 0; k < NM; ++k)
        if (k < n) v[k] -= piv[k];
}

// One tile.  Every thread of the block must call it (it synchronises); `label` is the cluster of
// trajectory i (any value outside [0, n_c) = not counted), `i` a valid (clamped) trajectory index.
// piv: [z pivot (at 0) | x pivot (at MTM_MAX_DIM)], kernel-parameter space (uniform loads).
// PREFETCH (stand-alone statistics kernel only): the rows come from HBM and every load below is consumed where it
// is issued — ask for all 2T rows of the trajectory up front.  In the fused E-step the rows are in L2 / L1 already
// (the same prefetches there measured 0.2 % slower) and its code must not change shape, so it is a template flag.
#ifdef MTM_STATS_TIMING  // diagnostic variant (tools/build_variant.sh): cycles per phase of block 0, printed by estep.cu
__device__ long long g_stats_cyc[8];
#define MTM_TICK(k)                                                                \
    do {                                                                           \
        const long long _now = clock64();                                          \
        if (blockIdx.x == 0 && threadIdx.x == 0) g_stats_cyc[k] += _now - _tick;   \
        _tick = _now;                                                              \
    } while (0)
#else
#define MTM_TICK(k)
#endif
template <class DM, bool PREFETCH = false>
__device__ __forceinline__ void stats_tile(const DM &dm, const double *__restrict__ z, const double *__restrict__ x,
                                           int64_t ld, int T, int64_t i, int label, int n_c, unsigned char *smem_base,
                                           const double *piv) {
    // carved here, per call, from ONE base pointer: keeping five derived pointers alive across the caller's
    // E-step loop costs uniform registers that the parameter operands of that loop need
    StatsSmem sm;
    sm.carve(smem_base, n_c, dm.d(), dm.l());
    const int d = dm.d(), l = dm.l();
    constexpr int DMX = DM::DM, LMX = DM::LM;
    constexpr int K1 = DMX, K2 = 2 * DMX, K3 = DMX + LMX;
    constexpr int F1M = Moments<K1>::F, F2M = Moments<K2>::F, F3M = Moments<K3>::F;
    const int F1 = 1 + d + d * (d + 1) / 2, F2 = 1 + 2 * d + d * (2 * d + 1);
    const int F3 = 1 + (d + l) + (d + l) * (d + l + 1) / 2;
    const int sstride = 1 + F1 + F2 + F3;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = kSThreads / 32;
    const int64_t zs = ld * d, xs = ld * l;
    double *s_acc = sm.acc, *s_feat = sm.feat;
    int *s_start = sm.start, *s_wcnt = sm.wcnt;
    if (label < 0 || label >= n_c) label = -1;
#ifdef MTM_STATS_TIMING
    long long _tick = clock64();
#endif

    // ---- stable counting sort of the tile's labels -> member lists ----
    int my_rank = 0;
    for (int c = 0; c < n_c; ++c) {
        const unsigned int m = __ballot_sync(0xffffffffu, label == c);
        if (lane == 0) s_wcnt[warp * n_c + c] = __popc(m);
        if (label == c) my_rank = __popc(m & ((1u << lane) - 1u));
    }
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int c = 0; c < n_c; ++c) {
            s_start[c] = run;
            for (int w = 0; w < NW; ++w) run += s_wcnt[w * n_c + c];
        }
        s_start[n_c] = run;
    }
    __syncthreads();
    int pos = -1;  // position of this trajectory in the cluster-sorted order of the tile
    if (label >= 0) {
        int off = s_start[label];
        for (int w = 0; w < warp; ++w) off += s_wcnt[w * n_c + label];
        pos = off + my_rank;
    }
    // cluster sizes: exact integers, carried as doubles (< 2^53)
    for (int c = tid; c < n_c; c += kSThreads) {
        double *acc = s_acc + 2 * ((size_t)c * sstride);
        dd_add(acc[0], acc[1], (double)(s_start[c + 1] - s_start[c]));
    }

    MTM_TICK(0);  // sort
    const double *z_i = z + i * d;
    const double *x_i = x + i * l;
    if (PREFETCH) {
        for (int t = 0; t < T; ++t) {
            asm volatile("prefetch.global.L1 [%0];" ::"l"(z_i + (int64_t)t * zs));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(x_i + (int64_t)t * xs));
        }
    }

    // NOTE: the two passes deliberately load each step right before its use.  A one-step-ahead software
    // prefetch here (tried) made ptxas 12.9 drop the uniform-datapath parameter operands of the E-step time
    // loop this code is inlined behind (tools/sass_reuse.py: 0.89 -> 0.79), a net loss.
    // ---- pass 1: initial-state and transition moments ----
    {
        double f1[F1M], f2[F2M];
        MTM_UNROLL
        for (int k = 0; k < F1M; ++k) f1[k] = 0.0;
        MTM_UNROLL
        for (int k = 0; k < F2M; ++k) f2[k] = 0.0;
        double zc[DMX], zn[DMX];
        load_row<DMX>(z_i, zc, d, DM::kStatic);
        centre_row<DMX>(zc, d, piv);
        bool okc = all_finite<DMX>(zc, d);
        add_moments<K1>(f1, zc, d, okc && label >= 0);
        for (int t = 0; t + 1 < T; ++t) {
            load_row<DMX>(z_i + (int64_t)(t + 1) * zs, zn, d, DM::kStatic);
            centre_row<DMX>(zn, d, piv);
            const bool okn = all_finite<DMX>(zn, d);
            double u[K2];
            MTM_UNROLL
            for (int k = 0; k < DMX; ++k)
                if (k < d) {
                    u[k] = zc[k];
                    u[d + k] = zn[k];
                }
            add_moments<K2>(f2, u, 2 * d, okc && okn && label >= 0);
            MTM_UNROLL
            for (int k = 0; k < DMX; ++k) zc[k] = zn[k];
            okc = okn;
        }
        MTM_TICK(1);  // pass 1
        reduce_by_cluster<F1M>(f1, F1, s_feat, pos, s_start, n_c, s_acc, sstride, 1);
        MTM_TICK(2);
        reduce_by_cluster<F2M>(f2, F2, s_feat, pos, s_start, n_c, s_acc, sstride, 1 + F1);
        MTM_TICK(3);
    }
    // ---- pass 2: measurement moments ----
    {
        double f3[F3M];
        MTM_UNROLL
        for (int k = 0; k < F3M; ++k) f3[k] = 0.0;
        // the rows of two time steps are requested together: half as many dependent L2 latencies (-0.5 % on the fused
        // kernel; the same batching in pass 1 costs 7 registers and, with them, the uniform-datapath form of the E-step
        // time loop this code is inlined behind — checked in SASS)
        for (int t = 0; t < T; t += 2) {
            double zc[2][DMX], xc[2][LMX];
            const bool second = t + 1 < T;  // block-uniform
            MTM_UNROLL
            for (int s = 0; s < 2; ++s) {
                const int ts = (s == 1 && !second) ? t : t + s;
                load_row<DMX>(z_i + (int64_t)ts * zs, zc[s], d, DM::kStatic);
                load_row<LMX>(x_i + (int64_t)ts * xs, xc[s], l, DM::kStatic);
            }
            MTM_UNROLL
            for (int s = 0; s < 2; ++s) {
                centre_row<DMX>(zc[s], d, piv);
                centre_row<LMX>(xc[s], l, piv + MTM_MAX_DIM);
                const bool ok = all_finite<DMX>(zc[s], d) && all_finite<LMX>(xc[s], l) && (s == 0 || second);
                double w[K3];
                MTM_UNROLL
                for (int k = 0; k < DMX; ++k)
                    if (k < d) w[k] = zc[s][k];
                MTM_UNROLL
                for (int k = 0; k < LMX; ++k)
                    if (k < l) w[d + k] = xc[s][k];
                add_moments<K3>(f3, w, d + l, ok && label >= 0);
            }
        }
        MTM_TICK(4);  // pass 2
        reduce_by_cluster<F3M>(f3, F3, s_feat, pos, s_start, n_c, s_acc, sstride, 1 + F1 + F2);
        MTM_TICK(5);
    }
}

}  // namespace mtm
