// Shared declarations of the mtm_b200 library (internal; the public surface is include/mtm_b200.h).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "../../include/mtm_b200.h"

namespace mtm {

#ifndef MTM_E_THREADS
#define MTM_E_THREADS 128  // tuning knob of tools/build_variant.sh (64 and 96 measured slower, profiles/r2_experiments.md)
#endif
constexpr int kEThreads = MTM_E_THREADS;  // E-step: one thread per trajectory, 4 warps per block
constexpr int kSThreads = MTM_E_THREADS;  // statistics: one thread per trajectory, tile = block

// Device-resident control block of the EM loop: written by the kernels, polled by the host.
struct Ctrl {
    int32_t done;       // loop finished: every later kernel of the enqueued batch returns at once
    int32_t exit_code;  // MTM_EXIT_*
    int32_t n_e_steps;
    int32_t n_m_steps;
    int32_t status;     // MTM_STATUS_* bits
    int32_t pad;
    unsigned long long n_switches;  // accumulated by the E-step kernel, consumed by finalize
    long long last_switches;
    long long min_count;
};

// Delta statistics: the reduced (double-double) statistics of the last two label vectors this handle ran a fused
// E-step on, keyed by a 128-bit hash of the labels.  When the labels an E-step starts from are one of them and few
// trajectories changed cluster in the iteration before, the fused kernel only visits the trajectories that switch
// (stats_tile.cuh: stats_delta_warp) and the reduction adds their moments to / subtracts them from the cached sums.
// Everything is decided on the device (label_hash_kernel, misc.cu): no host round trip, and labels the caller
// changed behind the library's back simply miss the cache.
struct StatsCacheEntry {
    unsigned long long h0, h1;  // label hash
    long long n;                // trajectories hashed
    long long sw;               // switches of the E-step that produced these labels
    unsigned long long stamp;   // age (replacement)
    int n_c;
    int valid;
};
struct StatsCache {
    StatsCacheEntry e[2];
    unsigned long long acc0, acc1;  // hash pass in flight (zero between passes)
    unsigned long long clock;
    unsigned int ticket;            // last-block detection of the hash pass
    int mode;                       // this iteration: 0 = full statistics, 1 = delta against entry `src`
    int src, dst;                   // entry read / entry written by this iteration's reduction
    unsigned int next_chunk;        // work counter of estep_delta_kernel (32-trajectory chunks), zeroed by the pre pass
};

struct EStepArgs {
    const double *z;  // (T, n, d) or nullptr
    const double *x;  // (T, n, l)
    int64_t n;
    int64_t ld;  // trajectories per time step in memory (>= n): step stride is ld*d / ld*l doubles
    int T, T0, d, l;
    const double *params;  // (n_c, stride)
    int n_c;
    int32_t *assign;       // nullable
    int count_switches;
    double *ll, *probs, *prenorm;  // nullable (n_c, n)
    Ctrl *ctrl;                    // never null
    unsigned long long *counts;    // (n_c) cluster sizes after the E-step, nullable
    double *score_partials;        // nullable: (grid, 2) block sums of Q and LL terms
    int check_done;                // honour ctrl->done (EM loop) or not (stand-alone calls)
    int R;                         // batched restarts: gridDim.y independent fits (0/1 = single fit); arrays
                                   // params (R, n_c, stride), assign (R, n), ctrl [R], counts (R, MAX_CLUSTERS),
                                   // score_partials (R, grid, 2), stats_partials (R, grid, n_c, sstride, 2)
    int tshift;                    // always 0 (see filter_eval): a never-taken uniform branch that fixes the code shape
    double *stats_partials;        // nullable: fused mode — the M-step statistics of the NEW labels are
                                   // accumulated in the same pass: (grid, n_c, sstride, 2) block sums
    double pivot[2 * MTM_MAX_DIM]; // fused mode: the moments are taken about this point: z - pivot[0..d), x - pivot[8..8+l)
    StatsCache *cache;             // nullable: fused mode may run as delta statistics (cache->mode, block-uniform)
    int delta_grid;                // blocks of the delta twin
};

struct StatsArgs {
    const double *z, *x;
    int64_t n;
    int64_t ld;  // trajectories per time step in memory (>= n)
    int T, d, l;
    const int32_t *assign;
    int n_c;
    double *partials;  // (R, grid, n_c, sstride, 2): block sums as (hi, lo) double-double pairs
    const Ctrl *ctrl;  // [R]
    int check_done;
    int R;             // batched restarts (gridDim.y); assign is (R, n)
    double pivot[2 * MTM_MAX_DIM];  // moments are taken about this point: z - pivot[0..d), x - pivot[8..8+l)
};

struct FinalizeArgs {
    const double *stats;  // (n_c, sstride)
    int n_c, d, l;
    int64_t n_total;
    double alpha;
    double *params;  // (n_c, pstride) out
    Ctrl *ctrl;
    int mode;  // 0: plain M-step (always write); 1: EM-loop step (decide converged / guard / continue)
    int n_steps;
    long long *switch_hist;  // nullable; (R, n_steps)
    int R;                   // batched restarts: one block per restart; stats (R, n_vals + kTail), params (R, n_c, pstride)
    double pivot[2 * MTM_MAX_DIM];  // the point the moments in `stats` were taken about (zeros = raw moments)
    // variant 1 = framework_extended semantics (mtm_mstep_ext / mtm_em_fit_ext): nanmean initial mean from
    // `nanmean` (n_c, 2d) [sum | count per coordinate], 1e-6 ridge when alpha == 0, the loop rules of SSMM.fit
    int variant;
    const double *nanmean;
};

// peer-memory mailboxes of the one-shot all-reduce (p2p.cu); at most 16 ranks (one NVSwitch box)
constexpr int kMaxPeers = 16;
struct P2PView {
    double *data[kMaxPeers];         // data[q]: rank q's mailbox, mapped here: [2][world][cap] doubles
    unsigned int *flags[kMaxPeers];  // flags[q]: rank q's flags: [2][world]
    int world, rank, cap;
    unsigned long long timeout_ns;   // how long a gather waits for its peers before MTM_STATUS_COMM_TIMEOUT
};
cudaError_t launch_p2p_allreduce(double *buf, int count, const P2PView &v, unsigned int seq, unsigned int *done_counter,
                                 Ctrl *ctrl, int R, cudaStream_t s);

// launchers (one translation unit per kernel family so that nvcc can build them in parallel)
cudaError_t launch_estep(const EStepArgs &a, int grid, size_t smem, cudaStream_t s);
cudaError_t launch_estep_prefix(const EStepArgs &a, double *pre, int grid, cudaStream_t s);  // pre: (T0, n_c, n)
int estep_grid(int64_t n, int sm_count, bool fused);
size_t estep_smem(int n_c, int T0, int d, int l, bool fused);

cudaError_t launch_stats(const StatsArgs &a, int grid, cudaStream_t s);
int stats_grid(int64_t n, int sm_count);
size_t stats_smem(int n_c, int d, int l);
// out (n_vals + kTail): reduced statistics, then the tail that rides through the all-reduce with them:
//   [n_switches of the preceding E-step | local n | 1 if this rank's E-step raised NONFINITE | ... SINGULAR]
// (all exact small integers as doubles).  After the all-reduce every rank sees the same switch count, the same
// total n and the same failure bits, so the control flow of finalize_kernel cannot diverge between ranks.
// `scratch` holds kReduceSlices * n_vals * 2 doubles (two-stage fixed-order reduction).
constexpr int kReduceSlices = 64;
constexpr int kTail = 4;
// cache / bases (nullable, R == 1 only): cache->mode == 0: bases[dst] = the reduced (hi, lo) sums; == 1: the block
// partials are DIFFERENCES (from grid_delta blocks), bases[dst] = bases[src] + them; out = bases[dst] rounded.
// bases: 2 x n_vals x 2 doubles.
cudaError_t launch_stats_reduce(const double *partials, int grid, int n_vals, double *out, double *scratch,
                                const Ctrl *ctrl, int check_done, int with_switches, double n_local, int R,
                                cudaStream_t s, const StatsCache *cache = nullptr, double *bases = nullptr,
                                int grid_delta = 0);
// blocks of estep_delta_kernel: one wave (the kernel hands out 32-trajectory chunks from a counter)
int estep_delta_grid(int64_t n, int sm_count);
// 128-bit hash of a label vector + the cache decision of one fused iteration (misc.cu).  role 0 (before the E-step,
// on the labels it starts from): look the hash up, set cache->mode / src / dst (delta only if allow_delta and the
// matching entry's switch count is at most n / 8); role 1 (after it, on the new labels): publish entry dst.
cudaError_t launch_label_hash(const int32_t *assign, int64_t n, int n_c, StatsCache *cache, const Ctrl *ctrl,
                              int check_done, int role, int allow_delta, int sm_count, cudaStream_t s);
// shared memory the delta mode of the fused E-step needs inside the statistics region
size_t stats_delta_smem_bytes(int n_c, int d, int l);
// tail (nullable) = the kTail trailing slots of the all-reduced statistics buffer
cudaError_t launch_finalize(const FinalizeArgs &a, const double *tail, cudaStream_t s);
// per-feature [sum | count] of the finite entries of a (T, n, w) array: out[0..8) sums, out[8..16) counts;
// scratch: pivot_scratch_doubles() doubles
size_t pivot_scratch_doubles(int sm_count);
cudaError_t launch_pivot_sums(const double *a, int64_t n, int64_t ld, int T, int w, int sm_count, double *scratch,
                              double *out16, cudaStream_t s);
// per-cluster, per-coordinate [sum | count] of the non-NaN entries of z_1 (np.nanmean, SSLG:43-44): out (n_c, 2d);
// scratch: nanmean_scratch_doubles() doubles
size_t nanmean_scratch_doubles(int n_c, int d, int sm_count);
cudaError_t launch_nanmean(const double *z, int64_t n, int d, const int32_t *assign, int n_c, int sm_count,
                           double *scratch, double *out, const Ctrl *ctrl, int check_done, cudaStream_t s);
cudaError_t launch_copy_labels(int32_t *dst, const int32_t *src, int64_t n, const Ctrl *ctrl, int check_done,
                               cudaStream_t s);
cudaError_t launch_counts(const int32_t *assign, int64_t n, int n_c, unsigned long long *counts, cudaStream_t s);
cudaError_t launch_score_reduce(const double *partials, int grid, double *out2, int R, cudaStream_t s);
cudaError_t run_fp64_peak(int reps, int sm_count, cudaStream_t s, double *tflops);
// device twin of numpy's default_rng(seed).integers(0, n_c, n); st = {state_hi, state_lo, inc_hi, inc_lo}
cudaError_t launch_rng_reject(const unsigned long long st[4], unsigned int n_c, long long n_draws, long long *rej_list,
                              int rej_cap, int *rej_count, cudaStream_t s);
cudaError_t launch_rng_emit(const unsigned long long st[4], unsigned int n_c, long long n, const long long *rej_sorted,
                            int n_rej, int32_t *out, cudaStream_t s);

// pattern-ordered residency (order.cu)
size_t order_scratch_bytes(int64_t n, int sm_count);
cudaError_t launch_pattern_order(const double *z, int64_t n, int64_t ld, int T, int d, int sm_count, void *scratch,
                                 int64_t *perm, cudaStream_t s);
cudaError_t launch_permute_f64(const double *src, double *dst, int64_t rows, int64_t n, int w, const int64_t *perm,
                               int inverse, cudaStream_t s);
cudaError_t launch_permute_i32(const int32_t *src, int32_t *dst, int64_t rows, int64_t n, const int64_t *perm,
                               int inverse, cudaStream_t s);

// k-means initial assignment (kmeans.cu)
cudaError_t launch_first_rows(const double *z, int64_t n, int64_t ld, int T, int d, double *rows, int sm_count,
                              cudaStream_t s);
size_t kmeans_scratch_doubles(int64_t n, int w, int n_c, int sm_count);
cudaError_t run_kmeans(double *rows, int64_t n, int w, int n_c, int64_t first_center, const double *u_host, int n_trials,
                       int max_iter, double rel_tol, int32_t *labels, double *centers, double *scratch, int sm_count,
                       cudaStream_t s, int *n_iter_out, int *how_out, int *empty_out, double *tol_out, int64_t *launches);

// 128-bit order-independent key of a (T, n, w) array (misc.cu); out2 must be zeroed by the caller
cudaError_t launch_fingerprint(const double *a, int64_t T, int64_t n, int w, const int64_t *perm,
                               unsigned long long seed, unsigned long long *out2, int sm_count, cudaStream_t s);

__host__ __device__ inline int param_stride(int d, int l) { return 1 + d + 3 * d * d + d * l + l * l; }
// [N | n0, sum z (d), sum zz' (d(d+1)/2) | n1, sum u (2d), sum uu' (d(2d+1)) | n2, sum w (d+l), sum ww' ((d+l)(d+l+1)/2)]
__host__ __device__ inline int stats_stride(int d, int l) {
    const int D = d + l;
    return 1 + (1 + d + d * (d + 1) / 2) + (1 + 2 * d + d * (2 * d + 1)) + (1 + D + D * (D + 1) / 2);
}

}  // namespace mtm
