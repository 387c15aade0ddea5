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

// ---- delta statistics ---------------------------------------------------------------------------------------------
// Late EM iterations move few trajectories (1.4 % in the sixth iteration of the 10 M benchmark fit), and the
// sufficient statistics of the new labels are those of the old ones plus / minus the moments of the movers.  In delta
// mode (StatsCache, common.cuh) every WARP handles its own movers, one at a time, with the lanes spread over the
// entries of a cluster's statistics block — no block-wide barrier, no staging, nothing at all for a warp without a
// mover (64 % of the warps at 1.4 %).  A mover's per-trajectory moments are formed with the very operations of
// add_moments above (fma over ascending t), so they are bit-identical to what the full pass adds for it; only the
// order in which trajectories are summed differs, as it does between grids or GPU counts.
//
// Step vector of time step t across the lanes: [0, d) z_t - pivot | [8, 8 + l) x_t - pivot | [16, 16 + d) z_{t+1} -
// pivot; a table entry names the moment group and the two lanes whose product feeds it (lane 32 = the constant 1).
__host__ __device__ inline int delta_entry(int e, int d, int l) {
    if (e == 0) return 0;  // cluster size: handled by the caller
    int o = e - 1;
    for (int g = 1; g <= 3; ++g) {
        const int K = g == 1 ? d : (g == 2 ? 2 * d : d + l);
        const int F = 1 + K + K * (K + 1) / 2;
        if (o < F) {
            int a = -1, b = -1;  // -1: the constant 1 (count and plain sums)
            if (o >= 1 && o <= K) {
                a = o - 1;
            } else if (o > K) {
                int p = o - 1 - K;
                a = 0;
                while (p >= K - a) {
                    p -= K - a;
                    ++a;
                }
                b = a + p;
            }
            int la = 32, lb = 32;
            if (a >= 0) la = a < d ? a : (g == 2 ? 16 + (a - d) : 8 + (a - d));
            if (b >= 0) lb = b < d ? b : (g == 2 ? 16 + (b - d) : 8 + (b - d));
            return g | (la << 4) | (lb << 10);
        }
        o -= F;
    }
    return 0;
}

// shared memory of the delta mode, carved from the same region as StatsSmem: one (hi, lo) accumulator block per warp
// and the entry table.  Double-double although the differences are sums over a handful of trajectories: the delta
// kernel hands its chunks out dynamically, so WHICH warp adds a mover varies from run to run, and with ~106-bit sums
// the rounded result does not (short of a sum landing within 2^-106 of a rounding boundary).
__host__ __device__ inline size_t stats_delta_bytes(int n_c, int d, int l) {
    return (size_t)(kSThreads / 32) * n_c * stats_stride(d, l) * 16 + (size_t)stats_stride(d, l) * 4 + 16;
}

// One chunk of 32 trajectories; `w_acc` = this warp's accumulators (n_c x sstride x (hi, lo)), `s_ent` the entry table.
// Per mover: the rows of eight time steps are requested together (they come from L2: one latency per eight steps,
// not one per step — the first version, a load per step and per round of entries, spent 17 k cycles per mover), then
// every lane feeds its entries (one per round of 32) from the step vectors by shuffles.
template <class DM>
__device__ __forceinline__ void stats_delta_warp(const DM &dm, const double *__restrict__ z, const double *__restrict__ x,
                                                 int64_t ld, int T, int64_t i, int old_label, int new_label, bool valid,
                                                 int n_c, double *w_acc, const int *s_ent, const double *piv) {
    const unsigned int full = 0xffffffffu;
    unsigned int movers = __ballot_sync(full, valid && old_label != new_label);
    if (movers == 0) return;
    // rounds of 32 entries per statistics block (common.cuh stats_stride, as a constant expression)
    constexpr int DD = DM::DM, DL = DM::DM + DM::LM;
    constexpr int RMAX = (1 + (1 + DD + DD * (DD + 1) / 2) + (1 + 2 * DD + DD * (2 * DD + 1)) + (1 + DL + DL * (DL + 1) / 2) + 31) / 32;
    constexpr int CH = 8;  // time steps in flight
    const int lane = threadIdx.x & 31;
    const int d = dm.d(), l = dm.l();
    const int sstride = stats_stride(d, l);
    const int64_t zs = ld * d, xs = ld * l;
    const int sub = lane & 7, grp = lane >> 3;  // 0: z_t, 1: x_t, 2: z_{t+1}
    const bool has = (grp == 0 && sub < d) || (grp == 1 && sub < l) || (grp == 2 && sub < d);
    const double pv = has ? piv[(grp == 1 ? MTM_MAX_DIM : 0) + sub] : 0.0;
    const unsigned int dmask = (1u << d) - 1u, lmask = (1u << l) - 1u;
    const int shift = grp == 2 ? 1 : 0;
    const int64_t step = grp == 1 ? xs : zs;
    int ent[RMAX];
    MTM_UNROLL
    for (int r = 0; r < RMAX; ++r) ent[r] = (r * 32 + lane < sstride) ? s_ent[r * 32 + lane] : 0;
    while (movers) {
        const int src = __ffs(movers) - 1;
        movers &= movers - 1;
        const int64_t j = __shfl_sync(full, i, src);
        int lo = __shfl_sync(full, old_label, src);
        const int ln = __shfl_sync(full, new_label, src);
        if (lo < 0 || lo >= n_c) lo = -1;  // was not counted
        const double *row = grp == 1 ? x + j * l + sub : z + j * d + sub;
        double s_r[RMAX];
        MTM_UNROLL
        for (int r = 0; r < RMAX; ++r) s_r[r] = 0.0;
        for (int t0 = 0; t0 < T; t0 += CH) {
            double v[CH];
            MTM_UNROLL
            for (int k = 0; k < CH; ++k) {
                const int t = t0 + k + shift;
                v[k] = __longlong_as_double(0x7ff8000000000000LL);
                if (has && t < T) v[k] = __ldg(row + (int64_t)t * step) - pv;
            }
            MTM_UNROLL
            for (int k = 0; k < CH; ++k) {
                if (t0 + k < T) {  // warp-uniform
                    const unsigned int fin = __ballot_sync(full, finite_bits(v[k]));
                    const bool okz = (fin & dmask) == dmask, okx = ((fin >> 8) & lmask) == lmask,
                               okn = ((fin >> 16) & dmask) == dmask;
                    const bool first = t0 + k == 0;
                    MTM_UNROLL
                    for (int r = 0; r < RMAX; ++r) {
                        if (r * 32 < sstride) {  // warp-uniform
                            const int g = ent[r] & 15, la = (ent[r] >> 4) & 63, lb = (ent[r] >> 10) & 63;
                            const double va = __shfl_sync(full, v[k], la & 31), vb = __shfl_sync(full, v[k], lb & 31);
                            const bool cond = g == 1 ? (first && okz) : (g == 2 ? (okz && okn) : (g == 3 && okz && okx));
                            if (cond) s_r[r] = fma(la == 32 ? 1.0 : va, lb == 32 ? 1.0 : vb, s_r[r]);
                        }
                    }
                }
            }
        }
        MTM_UNROLL
        for (int r = 0; r < RMAX; ++r) {
            const int e = r * 32 + lane;
            if (e < sstride) {
                const double s = e == 0 ? 1.0 : s_r[r];
                double *an = w_acc + 2 * (ln * sstride + e);
                dd_add(an[0], an[1], s);
                if (lo >= 0) {
                    double *ao = w_acc + 2 * (lo * sstride + e);
                    dd_add(ao[0], ao[1], -s);
                }
            }
        }
        __syncwarp();
    }
}

}  // namespace mtm
