// Kernel 1 (joint) and kernel 3 (measurement-only): E-step of the hard-assignment EM.
//
// One thread per trajectory, looping over the clusters; per cluster the thread runs the
// register-resident sequential filter of filter.cuh, then forms L_c = pi_c * exp(ll_c) in the
// linear domain and keeps the running argmax with numpy's rules (first maximum wins, the first
// NaN wins), exactly like mle_cluster_assignment (framework/marginalizable_mixture_model.py:
// 760-771).  The epilogue optionally writes ll / probs / prenorm (:775-791), counts assignment
// switches (E_step, :1235-1252) and cluster sizes, and accumulates the block sums of
// Q and LL terms (:579-666) in a fixed order.
//
// Memory: for a fixed time step the 32 lanes of a warp read 32 consecutive trajectories, i.e.
// 32*d*8 contiguous bytes of z and 32*l*8 of x (16-byte vector loads, read-only path).  A
// trajectory is streamed from HBM once per E-step: the n_c-1 later cluster passes of the same
// thread hit L1/L2.  Parameters (n_c x 85 doubles at d=l=4) sit in the constant bank and reach the
// DFMAs as uniform-register operands (CONSTP; shared memory when they do not fit or for run-time
// dimensions).  With the resident set in pattern order (order.cu) whole warps take the point-mass
// shortcut of filter.cuh in the steps whose hidden state is completely present.
#include <cstdlib>

#include "common.cuh"
#define MTM_DEFINE_CONST_PARAMS
#include "filter.cuh"
#include "stats_tile.cuh"

namespace mtm {

// NOTE: max-threads-only launch bounds on purpose.  With a min-blocks argument ptxas (12.9) keeps the
// model parameters in vector registers (LDC R, c[3][R+imm]); without it they go through the uniform
// datapath (LDCU.64 UR, c[3][UR+imm] -> DFMA R, R, UR, R), which is what lifts the register-file
// read-bandwidth limit on the DFMA stream (profiles/r1a_dfma_probe.txt).
#ifdef MTM_E_MINBLOCKS
#define MTM_E_BOUNDS __launch_bounds__(kEThreads, MTM_E_MINBLOCKS)
#else
#define MTM_E_BOUNDS __launch_bounds__(kEThreads)
#endif

// parameter source of cluster c: constant bank (CONSTP) or the block's shared-memory copy
template <bool CONSTP>
struct PSel;
template <>
struct PSel<true> {
    using type = ConstSrc;
    static __device__ __forceinline__ ConstSrc get(const double *, int cg, int, int stride) { return ConstSrc{cg * stride}; }
};
template <>
struct PSel<false> {
    using type = SmemSrc;
    static __device__ __forceinline__ SmemSrc get(const double *s_par, int, int c, int stride) {
        return SmemSrc{s_par + c * stride};
    }
};

// FUSE: the sufficient statistics of the M-step (stats_tile.cuh) are accumulated for the labels
// this very pass assigns, while the tile's trajectories are still hot in L1/L2 and the FP64 work of
// the other resident blocks hides the latency: one HBM pass per EM iteration instead of two.
//
// Fusion is a RUN-TIME switch (a.stats_partials != nullptr) on purpose: with it as a template flag the
// non-fused instantiation happened to fall back to vector-register parameter loads (ptxas 12.9
// heuristics, see the note at MTM_E_BOUNDS); one code shape for both keeps the uniform-datapath form.
template <int D_, int L_, bool HAS_Z, bool CONSTP>
__global__ void MTM_E_BOUNDS estep_kernel(const EStepArgs a) {
    using DM = Dims<D_, L_>;
    using PS = PSel<CONSTP>;
    const DM dm{a.d, a.l};
    // restart dimension (batched restarts, K4): blockIdx.y selects one independent fit — its own
    // parameters, labels, control block and partial sums; gridDim.y = 1 everywhere else
    const int r = blockIdx.y;
    Ctrl *const ctrl = a.ctrl + r;
    if (a.check_done && ctrl->done) return;
    if (a.cache != nullptr && a.cache->mode != 0) return;  // this iteration runs as delta statistics (estep_delta_kernel)
    const double *const params = a.params + (size_t)r * a.n_c * dm.stride();
    int32_t *const assign = a.assign ? a.assign + (size_t)r * a.n : nullptr;
    unsigned long long *const counts = a.counts ? a.counts + (size_t)r * MTM_MAX_CLUSTERS : nullptr;
    double *const score_partials = a.score_partials ? a.score_partials + 2 * (size_t)r * gridDim.x : nullptr;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int stride = dm.stride();
    double *s_par = reinterpret_cast<double *>(smem_raw);                 // n_c * stride (unused if CONSTP)
    double *s_logpi = s_par + (CONSTP ? 0 : a.n_c * stride);              // n_c
    double *s_pi = s_logpi + a.n_c;                                       // n_c
    uint32_t *s_fin = reinterpret_cast<uint32_t *>(s_pi + a.n_c);         // n_c * T0
    unsigned int *s_cnt = s_fin + a.n_c * a.T0;                           // n_c
    unsigned int *s_live = s_cnt + a.n_c;                                 // n_c: any coordinate of the cluster's mean finite
    const bool FUSE = a.stats_partials != nullptr;  // block-uniform
    unsigned char *const stats_base =
        smem_raw + (((reinterpret_cast<unsigned char *>(s_live + a.n_c) - smem_raw) + 15) & ~size_t(15));
    __shared__ double s_red[2][kEThreads / 32];
    __shared__ unsigned int s_sw;

    const int tid = threadIdx.x;
    if (!CONSTP)
        for (int k = tid; k < a.n_c * stride; k += kEThreads) s_par[k] = params[k];
    for (int k = tid; k < a.n_c; k += kEThreads) s_cnt[k] = 0;
    if (FUSE)
        for (int k = tid; k < 2 * a.n_c * stats_stride(dm.d(), dm.l()); k += kEThreads)
            reinterpret_cast<double *>(stats_base)[k] = 0.0;
    if (tid == 0) s_sw = 0;
    __syncthreads();
    // per cluster: pi, log pi, finiteness masks of the unconditional mean (8 lanes per cluster, 16 clusters per pass)
    for (int c0 = 0; c0 < a.n_c; c0 += kEThreads / 8) {
        const int c_raw = c0 + (tid >> 3);
        const bool c_ok = c_raw < a.n_c;
        const int c = c_ok ? c_raw : a.n_c - 1;
        const double *row = params + (size_t)c * stride;  // global memory on purpose, see mean_finite_masks_lanes
        if (c_ok && (tid & 7) == 0) {
            const double pi_c = __ldg(row);
            s_pi[c] = pi_c;
            s_logpi[c] = log(pi_c);
        }
        mean_finite_masks_lanes(dm, row, a.T0, s_fin + c * a.T0, c_ok);
    }
    __syncthreads();
    for (int c = tid; c < a.n_c; c += kEThreads) {
        uint32_t any = 0;
        for (int t = 0; t < a.T0; ++t) any |= s_fin[c * a.T0 + t];
        s_live[c] = any;
    }
    __syncthreads();

    const int64_t n = a.n;
    const int64_t n_tiles = (n + kEThreads - 1) / kEThreads;
    const int64_t zs = a.ld * dm.d(), xs = a.ld * dm.l();
    double q_acc = 0.0, l_acc = 0.0;
    unsigned int sw_acc = 0;
    int bad = 0;

#ifdef MTM_STATS_TIMING
    long long t_e = 0, t_all0 = clock64();
    int n_my_tiles = 0;
#endif
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
#ifdef MTM_STATS_TIMING
        const long long t_tile0 = clock64();
        ++n_my_tiles;
#endif
        const int64_t i_raw = tile * kEThreads + tid;
        const bool valid = i_raw < n;
        const int64_t i = valid ? i_raw : n - 1;  // clamp: the tail lanes compute and discard
        const double *z_i = HAS_Z ? a.z + i * dm.d() : nullptr;
        const double *x_i = a.x + i * dm.l();

        double best = 0.0, best_ll = 0.0, sum_l = 0.0;
        int arg = 0;
        for (int c = 0; c < a.n_c; ++c) {
            // A dead cluster — its unconditional mean is non-finite at every step (NaN m after an empty
            // initial-state sample) — has no present coordinate at all: the reference's sub-setting leaves an
            // empty block and ll = -0.0 (SSM:769-773).  It runs ZERO filter steps (block-uniform trip count; the
            // empty filter finishes with exactly -0.0): without this such a cluster defeats the point-mass vote
            // and drags every warp through the slow generic path for nothing.  (A trip count rather than a branch
            // around the call: the branch made ptxas 12.9 drop the uniform-datapath parameter operands of the
            // observation-only instantiation.)
            const double ll = filter_eval<DM, HAS_Z>(dm, z_i, zs, x_i, xs, s_live[c] ? a.T0 : 0,
                                                     PS::get(s_par, r * a.n_c + c, c, stride), s_fin + c * a.T0, bad,
                                                     a.tshift);
            const double lc = s_pi[c] * exp(ll);
            if (valid) {
                if (a.ll) a.ll[(int64_t)c * n + i] = ll;
                if (a.prenorm) a.prenorm[(int64_t)c * n + i] = s_logpi[c] + ll;
                if (a.probs) a.probs[(int64_t)c * n + i] = lc;
            }
            sum_l += lc;
            // np.argmax: strict > keeps the first maximum; a NaN beats everything, the first NaN stays
            const bool take = (c == 0) || (!(best != best) && (lc > best || lc != lc));
            if (take) {
                best = lc;
                best_ll = ll;
                arg = c;
            }
        }
        if (valid) {
            if (a.probs) {
                for (int c = 0; c < a.n_c; ++c) {
                    const int64_t o = (int64_t)c * n + i;
                    a.probs[o] = a.probs[o] / sum_l;
                }
            }
            if (assign) {
                if (a.count_switches) sw_acc += (assign[i] != arg) ? 1u : 0u;
                assign[i] = arg;
            }
            q_acc += s_logpi[arg] + best_ll;
            l_acc += log(sum_l);
        }
        if (counts) {
            for (int c = 0; c < a.n_c; ++c) {
                const unsigned int m = __ballot_sync(0xffffffffu, valid && arg == c);
                if ((tid & 31) == 0 && m) atomicAdd(&s_cnt[c], (unsigned int)__popc(m));
            }
        }
#ifdef MTM_STATS_TIMING
        t_e += clock64() - t_tile0;
#endif
        if (FUSE) stats_tile(dm, a.z, a.x, a.ld, a.T, i, valid ? arg : -1, a.n_c, stats_base, a.pivot);
    }
#ifdef MTM_STATS_TIMING
    if (FUSE && blockIdx.x == 0 && tid == 0 && n_my_tiles > 4)
        printf("block 0: %d tiles, cycles per tile: E+argmax %lld | sort %lld pass1 %lld red1 %lld red2 %lld pass2 %lld red3 %lld | total %lld\n",
               n_my_tiles, t_e / n_my_tiles, g_stats_cyc[0] / n_my_tiles, g_stats_cyc[1] / n_my_tiles, g_stats_cyc[2] / n_my_tiles,
               g_stats_cyc[3] / n_my_tiles, g_stats_cyc[4] / n_my_tiles, g_stats_cyc[5] / n_my_tiles,
               (clock64() - t_all0) / n_my_tiles);
#endif

    // block epilogue: integer counters by atomics (exact, order-free), fp64 sums in a fixed order
    if (bad) atomicOr(&ctrl->status, bad);
    if (a.count_switches) {
        MTM_UNROLL
        for (int o = 16; o > 0; o >>= 1) sw_acc += __shfl_xor_sync(0xffffffffu, sw_acc, o);
        if ((tid & 31) == 0 && sw_acc) atomicAdd(&s_sw, sw_acc);
    }
    if (score_partials) {
        MTM_UNROLL
        for (int o = 16; o > 0; o >>= 1) {
            q_acc += __shfl_xor_sync(0xffffffffu, q_acc, o);
            l_acc += __shfl_xor_sync(0xffffffffu, l_acc, o);
        }
        if ((tid & 31) == 0) {
            s_red[0][tid >> 5] = q_acc;
            s_red[1][tid >> 5] = l_acc;
        }
    }
    __syncthreads();
    if (tid == 0) {
        if (a.count_switches && s_sw) atomicAdd(&ctrl->n_switches, (unsigned long long)s_sw);
        if (score_partials) {
            double q = 0.0, l = 0.0;
            for (int w = 0; w < kEThreads / 32; ++w) {
                q += s_red[0][w];
                l += s_red[1][w];
            }
            score_partials[2 * blockIdx.x] = q;
            score_partials[2 * blockIdx.x + 1] = l;
        }
    }
    if (counts)
        for (int c = tid; c < a.n_c; c += kEThreads)
            if (s_cnt[c]) atomicAdd(&counts[c], (unsigned long long)s_cnt[c]);
    if (FUSE) {
        const int nv = 2 * a.n_c * stats_stride(dm.d(), dm.l());
        double *out = a.stats_partials + ((size_t)r * gridDim.x + blockIdx.x) * nv;
        for (int k = tid; k < nv; k += kEThreads) out[k] = reinterpret_cast<const double *>(stats_base)[k];
    }
}

// Delta-statistics twin of the fused kernel (common.cuh StatsCache, stats_tile.cuh stats_delta_warp): the same
// E-step and argmax, but instead of the tile-wide statistics pass every warp adds / subtracts the moments of its own
// trajectories that change cluster.  Launched right after estep_kernel on every fused iteration of a single fit;
// exactly one of the two does the work (cache->mode, set on the device by the label-hash pass), the other returns
// at once.  A separate kernel on purpose: any edit inside estep_kernel's tile loop is liable to cost it the
// uniform-datapath form of its parameter operands (see MTM_E_BOUNDS).
template <int D_, int L_, bool CONSTP>
__global__ void MTM_E_BOUNDS estep_delta_kernel(const EStepArgs a) {
    using DM = Dims<D_, L_>;
    using PS = PSel<CONSTP>;
    const DM dm{a.d, a.l};
    Ctrl *const ctrl = a.ctrl;
    if (a.check_done && ctrl->done) return;
    if (a.cache->mode == 0) return;  // full statistics this iteration (estep_kernel)
    const double *const params = a.params;
    int32_t *const assign = a.assign;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int stride = dm.stride();
    double *s_par = reinterpret_cast<double *>(smem_raw);                 // n_c * stride (unused if CONSTP)
    double *s_logpi = s_par + (CONSTP ? 0 : a.n_c * stride);              // n_c
    double *s_pi = s_logpi + a.n_c;                                       // n_c
    uint32_t *s_fin = reinterpret_cast<uint32_t *>(s_pi + a.n_c);         // n_c * T0
    unsigned int *s_live = s_fin + a.n_c * a.T0;                          // n_c
    double *const acc_all = reinterpret_cast<double *>(
        smem_raw + (((reinterpret_cast<unsigned char *>(s_live + a.n_c) - smem_raw) + 15) & ~size_t(15)));
    const int ss = stats_stride(dm.d(), dm.l());
    constexpr int NW = kEThreads / 32;
    const int nv = 2 * a.n_c * ss;  // (hi, lo) pairs of one accumulator block
    int *const s_ent = reinterpret_cast<int *>(acc_all + NW * nv);
    __shared__ unsigned int s_sw;

    const int tid = threadIdx.x, lane = tid & 31;
    if (!CONSTP)
        for (int k = tid; k < a.n_c * stride; k += kEThreads) s_par[k] = params[k];
    for (int k = tid; k < NW * nv; k += kEThreads) acc_all[k] = 0.0;
    for (int k = tid; k < ss; k += kEThreads) s_ent[k] = delta_entry(k, dm.d(), dm.l());
    if (tid == 0) s_sw = 0;
    __syncthreads();
    for (int c0 = 0; c0 < a.n_c; c0 += kEThreads / 8) {
        const int c_raw = c0 + (tid >> 3);
        const bool c_ok = c_raw < a.n_c;
        const int c = c_ok ? c_raw : a.n_c - 1;
        const double *row = params + (size_t)c * stride;
        if (c_ok && (tid & 7) == 0) {
            const double pi_c = __ldg(row);
            s_pi[c] = pi_c;
            s_logpi[c] = log(pi_c);
        }
        mean_finite_masks_lanes(dm, row, a.T0, s_fin + c * a.T0, c_ok);
    }
    __syncthreads();
    for (int c = tid; c < a.n_c; c += kEThreads) {
        uint32_t any = 0;
        for (int t = 0; t < a.T0; ++t) any |= s_fin[c * a.T0 + t];
        s_live[c] = any;
    }
    __syncthreads();

    // Work distribution: every WARP draws 32-trajectory chunks from one counter (zeroed by the label-hash pass that
    // selected this kernel).  The tile loop of this kernel has no block-wide barrier, so warps run free; one wave of
    // blocks then carries the whole launch with one prologue each and no tail — with estep_kernel's static striding
    // a 1.25 M shard (one rank of eight) paid 8 % for 12 rounds of block prologues or, with few blocks, for imbalance.
    const int64_t n = a.n;
    const unsigned int n_chunks = (unsigned int)((n + 31) / 32);
    const int64_t zs = a.ld * dm.d(), xs = a.ld * dm.l();
    unsigned int sw_acc = 0;
    int bad = 0;
    double *const w_acc = acc_all + (tid >> 5) * nv;
    for (;;) {
        unsigned int chunk = 0;
        if (lane == 0) chunk = atomicAdd(&a.cache->next_chunk, 1u);
        chunk = __shfl_sync(0xffffffffu, chunk, 0);
        if (chunk >= n_chunks) break;
        const int64_t i_raw = (int64_t)chunk * 32 + lane;
        const bool valid = i_raw < n;
        const int64_t i = valid ? i_raw : n - 1;
        const double *z_i = a.z + i * dm.d();
        const double *x_i = a.x + i * dm.l();
        double best = 0.0;
        int arg = 0;
        for (int c = 0; c < a.n_c; ++c) {
            const double ll = filter_eval<DM, true>(dm, z_i, zs, x_i, xs, s_live[c] ? a.T0 : 0,
                                                    PS::get(s_par, c, c, stride), s_fin + c * a.T0, bad, a.tshift);
            const double lc = s_pi[c] * exp(ll);
            const bool take = (c == 0) || (!(best != best) && (lc > best || lc != lc));
            if (take) {
                best = lc;
                arg = c;
            }
        }
        int old_lab = arg;
        if (valid) {
            old_lab = assign[i];
            sw_acc += (old_lab != arg) ? 1u : 0u;
            assign[i] = arg;
        }
        stats_delta_warp(dm, a.z, a.x, a.ld, a.T, i, old_lab, arg, valid, a.n_c, w_acc, s_ent, a.pivot);
    }

    if (bad) atomicOr(&ctrl->status, bad);
    MTM_UNROLL
    for (int o = 16; o > 0; o >>= 1) sw_acc += __shfl_xor_sync(0xffffffffu, sw_acc, o);
    if (lane == 0 && sw_acc) atomicAdd(&s_sw, sw_acc);
    __syncthreads();
    if (tid == 0 && s_sw) atomicAdd(&ctrl->n_switches, (unsigned long long)s_sw);
    // the warps' differences, added in warp order; the reduction kernels are shared with the full pass
    double *out = a.stats_partials + (size_t)blockIdx.x * nv;
    for (int k = tid; k < nv / 2; k += kEThreads) {
        double hi = 0.0, lo = 0.0;
        for (int w = 0; w < NW; ++w) {
            dd_add(hi, lo, acc_all[w * nv + 2 * k]);
            lo += acc_all[w * nv + 2 * k + 1];
        }
        out[2 * k] = hi;
        out[2 * k + 1] = lo;
    }
}

// fixed-order sum of the per-block (Q, LL) partials
__global__ void score_reduce_kernel(const double *__restrict__ partials_all, int grid, double *out_all) {
    const double *partials = partials_all + 2 * (size_t)blockIdx.x * grid;  // one block per restart
    double *out2 = out_all + 2 * blockIdx.x;
    __shared__ double s[2][256];
    double q = 0.0, l = 0.0;
    for (int b = threadIdx.x; b < grid; b += 256) {
        q += partials[2 * b];
        l += partials[2 * b + 1];
    }
    s[0][threadIdx.x] = q;
    s[1][threadIdx.x] = l;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            s[0][threadIdx.x] += s[0][threadIdx.x + o];
            s[1][threadIdx.x] += s[1][threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out2[0] = s[0][0];
        out2[1] = s[1][0];
    }
}

size_t estep_smem(int n_c, int T0, int d, int l, bool fused) {
    const size_t e = (size_t)n_c * param_stride(d, l) * 8 + (size_t)n_c * 16 + (size_t)n_c * T0 * 4 + (size_t)n_c * 8 + 32;
    return e + (fused ? stats_smem_bytes(n_c, d, l) : 0);
}

size_t stats_delta_smem_bytes(int n_c, int d, int l) { return stats_delta_bytes(n_c, d, l); }

int estep_delta_grid(int64_t n, int sm_count) {
    // one wave of the 3 blocks per SM that fit (148 registers x 128 threads); any other grid works too, the chunks
    // are drawn from a counter
    const int64_t blocks = (n + kEThreads - 1) / kEThreads;
    const int64_t cap = (int64_t)sm_count * 3;
    return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

int estep_grid(int64_t n, int sm_count, bool fused) {
    const int64_t tiles = (n + kEThreads - 1) / kEThreads;
    // Static tile striding (uniform cost per tile), blocks = a multiple of the SM count and of the 3 resident
    // blocks per SM.  Plain E-step: 32 per SM (30 / 36 / 48 measure the same).  Fused: every block also carries
    // n_c*106 double-double accumulators that a reduction has to read back, which argues for few blocks — but many
    // short-lived blocks de-phase the statistics epilogues of the co-resident blocks: at 10 M x 8 on the
    // pattern-ordered set 12 / 18 / 24 / 36 / 48 / 72 blocks per SM give 18.53 / 18.35 / 18.19 / 18.09 / 18.10 /
    // 18.08 ms (first build: 3 / 6 / 12 / 24 / 64 -> 23.8 / 23.0 / 22.4 / 22.1 / 22.0, profiles/r1h_grid_sweep.txt).
    // 36 per SM (72 MB of partials at n_c = 8).
    int per_sm = (fused ? 36 : 32) * 128 / kEThreads;
    if (const char *e = getenv(fused ? "MTM_FUSED_BLOCKS_PER_SM" : "MTM_ESTEP_BLOCKS_PER_SM")) {  // tuning knob
        const int v = atoi(e);
        if (v > 0) per_sm = v;
    }
    const int64_t cap = (int64_t)sm_count * per_sm;
    return (int)(tiles < cap ? (tiles > 0 ? tiles : 1) : cap);
}

template <int D_, int L_, bool HAS_Z, bool CONSTP>
static cudaError_t launch_one(const EStepArgs &a, int grid, size_t smem, cudaStream_t s) {
    auto k = estep_kernel<D_, L_, HAS_Z, CONSTP>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    k<<<dim3(grid, a.R > 1 ? a.R : 1), kEThreads, smem, s>>>(a);
    return cudaGetLastError();
}

template <int D_, int L_, bool CONSTP>
static cudaError_t launch_delta(const EStepArgs &a, int grid, size_t smem, cudaStream_t s) {
    auto k = estep_delta_kernel<D_, L_, CONSTP>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    k<<<a.delta_grid, kEThreads, smem, s>>>(a);
    return cudaGetLastError();
}

template <int D_, int L_>
static cudaError_t launch_dl(const EStepArgs &a, int grid, size_t smem, cudaStream_t s) {
    const size_t n_par = (size_t)(a.R > 1 ? a.R : 1) * a.n_c * param_stride(a.d, a.l);
    const bool delta = a.cache != nullptr;  // the twin runs the iteration when the device-side cache says so
    if (D_ > 0 && n_par <= MTM_CONST_DOUBLES) {
        // stage the parameter rows in the constant bank (device -> device, stream ordered)
        cudaError_t e = cudaMemcpyToSymbolAsync(c_par, a.params, n_par * sizeof(double), 0,
                                                cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) return e;
        e = a.z ? launch_one<D_, L_, true, true>(a, grid, smem, s) : launch_one<D_, L_, false, true>(a, grid, smem, s);
        if (e != cudaSuccess || !delta) return e;
        return launch_delta<D_, L_, true>(a, grid, smem, s);
    }
    cudaError_t e = a.z ? launch_one<D_, L_, true, false>(a, grid, smem, s) : launch_one<D_, L_, false, false>(a, grid, smem, s);
    if (e != cudaSuccess || !delta) return e;
    return launch_delta<D_, L_, false>(a, grid, smem, s);
}

cudaError_t launch_estep(const EStepArgs &a, int grid, size_t smem, cudaStream_t s) {
#define MTM_CASE(DD, LL) \
    if (a.d == DD && a.l == LL) return launch_dl<DD, LL>(a, grid, smem, s);
    MTM_CASE(4, 4)
#ifndef MTM_ONLY_44
    MTM_CASE(2, 4)
    MTM_CASE(2, 3)
    MTM_CASE(1, 1)
    MTM_CASE(3, 3)
    MTM_CASE(2, 2)
    MTM_CASE(5, 3)
#endif
#undef MTM_CASE
#ifdef MTM_ONLY_44
    return cudaErrorInvalidValue;
#else
    return launch_dl<0, 0>(a, grid, smem, s);  // generic run-time dimensions
#endif
}

cudaError_t launch_score_reduce(const double *partials, int grid, double *out2, int R, cudaStream_t s) {
    score_reduce_kernel<<<R > 1 ? R : 1, 256, 0, s>>>(partials, grid, out2);
    return cudaGetLastError();
}

}  // namespace mtm
