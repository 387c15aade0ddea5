/*
 * mtm_b200 — C ABI of the B200-native hard-assignment EM hot path for mixtures of
 * linear-Gaussian state-space models.
 *
 * Drop-in boundary for the numerics of burkh4rt/Multimodal-Trajectory-Modeling's
 *   framework/marginalizable_mixture_model.py   (MMM)  class MMLinGaussSS_marginalizable
 *   framework/marginalizable_state_space_model.py (SSM) mm / CC / multivariate_normal_log_likelihood
 * Each entry point cites the reference interface it replaces.  The reference has no FFI of
 * its own (it is pure Python); the binding a maintainer would add is the ctypes stub shown
 * in INTEGRATION.md (and shipped in multimodal-trajectory-modeling_b200/_capi.py).
 *
 * Conventions
 *  - every pointer named *_dev is DEVICE memory owned by the caller (e.g. a torch tensor's
 *    data_ptr()); the library never frees or retains it beyond the call, except the data
 *    set bound with mtm_set_data, which must stay alive while bound;
 *  - trajectories: states z (T, n, d) and observations x (T, n, l), C-contiguous float64,
 *    time-major, missing = any non-finite value (NaN or +-inf)  [MMM:60-63, SSM:770];
 *  - parameters: (n_c, mtm_param_stride(d,l)) float64 rows
 *        [pi | m (d) | S (d,d) | A (d,d) | Gamma (d,d) | H (d,l) | Lambda (l,l)], row-major,
 *    i.e. the reference attributes cluster_propensities / init_state_means / init_state_covs /
 *    transition_matrices / transition_covs / measurement_matrices / measurement_covs
 *    [MMM:83-125];
 *  - assignments: int32 on the device (the reference holds int64 from np.argmax; the host
 *    layer widens);
 *  - every function returns 0 on success, <0 on error (message: mtm_last_error); numerical
 *    trouble is NOT an error code but a bit in mtm_fit_info.status / the status word, so
 *    the host can emulate the reference's exceptions [MMM:1439-1442, 1463-1464];
 *  - work is enqueued on the handle's CUDA stream; only calls that return host values
 *    (mtm_em_fit, mtm_mstep, mtm_score, mtm_estep when info != NULL, mtm_fp64_peak)
 *    synchronise that stream.  One handle per (process, device); calls on one handle must
 *    be serialised by the caller (the reference class is not re-entrant either).
 */
#ifndef MTM_B200_H
#define MTM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MTM_ABI_VERSION 1
#define MTM_MAX_DIM 8        /* largest supported d and l */
#define MTM_MAX_CLUSTERS 64  /* largest n_c of one model */

/* status bits (mtm_fit_info.status, mtm_estep_info.status) */
#define MTM_STATUS_NONFINITE 1   /* a covariance pivot was NaN/inf: the reference's det/solve raise
                                    LinAlgError("Array must not contain infs or NaNs") [SSM:771-773] */
#define MTM_STATUS_SINGULAR 2    /* an exactly zero pivot: the reference's solve raises "singular" */
#define MTM_STATUS_EMPTY_FIT 4   /* a regression had no usable rows: lstsq / Ridge raise [MMM:1194,1226] */
#define MTM_STATUS_CHOL_FAIL 8   /* normal equations not positive definite in the M-step */
#define MTM_STATUS_COMM_TIMEOUT 16 /* a peer never arrived in the peer-memory all-reduce */

/* exit codes of the EM loop [MMM:1310-1366] */
#define MTM_EXIT_RUNNING 0
#define MTM_EXIT_CONVERGED 1  /* an E-step changed no assignment            [MMM:1342-1345] */
#define MTM_EXIT_GUARD 2      /* min cluster size <= 3 after an E-step       [MMM:1346-1356] */
#define MTM_EXIT_EXHAUSTED 3  /* n_steps E-steps done                        [MMM:1340]      */
#define MTM_EXIT_GUARD0 4     /* min cluster size <= 3 before the first M-step: nothing ran [MMM:1328-1336] */
#define MTM_EXIT_ERROR 5      /* a numerical failure on which the reference raises (status != 0): the loop
                                 stops at that step so the host can raise the same exception */
#define MTM_EXIT_EMPTY 6      /* mtm_em_fit_ext only: an E-step left a cluster without members; the labels of
                                 that E-step are discarded (framework_extended/state_space_model_mixture.py:122-124) */

typedef struct mtm_handle mtm_handle;

/* a data set in device memory; z_dev == NULL means "all hidden states missing"
 * (the observation-only route, MMM:997-1048, 1125-1165) */
typedef struct {
    const double *z_dev;
    const double *x_dev;
    int64_t n;
    int32_t T;
    int32_t d;
    int32_t l;
    int32_t reserved;
    int64_t ld; /* trajectories per time step in memory: element (t, i, j) of z sits at z_dev[(t*ld + i)*d + j].
                   0 means ld = n (contiguous).  ld > n addresses a slab of trajectories inside a larger
                   (T, ld, d) buffer without repacking. */
} mtm_data;

typedef struct {
    int64_t n_switches;  /* assignments changed by this E-step (global if a communicator is attached
                            and the call all-reduces; mtm_estep itself is local) */
    int64_t min_count;   /* smallest cluster size after the E-step (local) */
    int32_t status;      /* MTM_STATUS_* bits */
    int32_t stats_mode;  /* mtm_em_step: how the M-step statistics of this iteration were formed: 0 = summed over
                            all trajectories, 1 = delta statistics (see mtm_set_delta_stats); 0 elsewhere */
} mtm_estep_info;

typedef struct {
    int32_t exit_code;   /* MTM_EXIT_* */
    int32_t n_e_steps;   /* E-steps executed; the reference's reported "i" is n_e_steps-1 on convergence */
    int32_t status;      /* MTM_STATUS_* bits accumulated over the fit */
    int32_t n_m_steps;   /* M-steps executed */
    int64_t last_switches;
    int64_t min_count;
} mtm_fit_info;

/* ---- lifecycle ------------------------------------------------------------------------- */
int mtm_abi_version(void);
/* stream: a cudaStream_t (NULL = legacy default stream) */
int mtm_create(int device, void *stream, mtm_handle **out);
int mtm_destroy(mtm_handle *h);
int mtm_set_stream(mtm_handle *h, void *stream);
/* message of the last error on this handle (h == NULL: last error of mtm_create) */
const char *mtm_last_error(mtm_handle *h);

/* ---- layout helpers ---------------------------------------------------------------------- */
int mtm_param_stride(int d, int l); /* 1 + d + 3 d^2 + d l + l^2   (85 at d=l=4) */
int mtm_stats_stride(int d, int l); /* sufficient statistics per cluster (106 at d=l=4), see DESIGN.md */

/* ---- resident training set  [MMM:45-81: ctor stores states/observations] ---------------- */
int mtm_set_data(mtm_handle *h, const mtm_data *data);

/* ---- E-step ---------------------------------------------------------------------------------
 * Replaces, in one launch: conditional_log_likelihoods[_first_T0_steps] for every cluster
 * (MMM:439-530 -> SSM mm/CC/multivariate_normal_log_likelihood, SSM:259-364, 728-773),
 * mle_cluster_assignment (MMM:724-791: linear-domain argmax of pi_c exp(ll_c), first max wins,
 * NaN wins; optional probs and prenormalised log-probs), E_step's switch count (MMM:1235-1252),
 * and — with data->z_dev == NULL — observed_condl_log_lik_first_T0_steps /
 * observations_mle_cluster_assignment (MMM:997-1165).
 *   data           NULL = the resident set
 *   T0             time cutoff, clamped to data->T   (MMM:473, 1026)
 *   assign_dev     (n) int32, nullable.  If count_switches != 0 it is read as the old assignment
 *                  first and info->n_switches counts the changes.
 *   ll/probs/prenorm_dev  (n_c, n) float64 outputs, each nullable
 *   info           host struct, nullable; when given the call synchronises the stream      */
int mtm_estep(mtm_handle *h, const mtm_data *data, const double *params_dev, int n_c, int T0,
              int32_t *assign_dev, int count_switches, double *ll_dev, double *probs_dev,
              double *prenorm_dev, mtm_estep_info *info);

/* ---- all prefix log-likelihoods in one pass ------------------------------------------------------
 * ll_prefix_dev (T0, n_c, n) float64 out: ll_prefix[t, c, i] = conditional_log_likelihoods_first_T0_steps(c, t + 1)[i]
 * (MMM:439-494; with data->z_dev == NULL observed_condl_log_lik_first_T0_steps, MMM:997-1048) for every
 * t < T0 at once.  This is what cluster_propensities_over_time / observed_cluster_propensities_over_time
 * (MMM:532-577, 1077-1123) need; the reference filters T0 (T0 + 1) / 2 steps per (trajectory, cluster) for
 * it, the sequential filter passes through every prefix anyway.
 * status_host (nullable): MTM_STATUS_* bits; when given the call synchronises. */
int mtm_estep_prefix(mtm_handle *h, const mtm_data *data, const double *params_dev, int n_c, int T0,
                     double *ll_prefix_dev, int32_t *status_host);

/* ---- scores -----------------------------------------------------------------------------------
 * out_host[0] = Q  = sum_i log pi_{a_i} + ll[a_i, i] with a = argmax   (e_complete_data_log_lik, MMM:579-624)
 * out_host[1] = LL = sum_i log sum_c pi_c exp(ll[c,i])                  (model_log_likelihood,   MMM:626-666)
 * Sums are all-reduced over the attached communicator.  Synchronises. */
int mtm_score(mtm_handle *h, const mtm_data *data, const double *params_dev, int n_c, double *out_host,
              int32_t *status_host);

/* ---- M-step -------------------------------------------------------------------------------------
 * mtm_mstats: per-cluster sufficient statistics of the resident set under assign_dev
 *   (restates M_step's masked gathers as moments, MMM:1254-1308); stats_dev (n_c, stats_stride),
 *   local to this rank (no collective).
 * mtm_finalize: parameters from (all-reduced) statistics: pi = N_c / n_total (MMM:1261-1263),
 *   m, S (np.cov, ddof=1; MMM:1267-1270), (A, Gamma) and (H, Lambda) by regress / regress_alpha
 *   (MMM:1167-1233): alpha <= 2 eps -> least squares, else ridge (X'X + alpha I) B = X'Y; residual
 *   covariance mean-subtracted with N-1.
 * mtm_mstep: mstats -> all-reduce over the attached communicator -> finalize; writes
 *   params_dev; info (nullable) receives min_count and status.  Synchronises if info != NULL. */
int mtm_mstats(mtm_handle *h, const int32_t *assign_dev, int n_c, double *stats_dev);
int mtm_finalize(mtm_handle *h, const double *stats_dev, int n_c, int d, int l, int64_t n_total,
                 double alpha, double *params_dev, int32_t *status_dev);
int mtm_mstep(mtm_handle *h, const int32_t *assign_dev, int n_c, double alpha, double *params_dev,
              mtm_estep_info *info);

/* ---- one EM iteration ------------------------------------------------------------------------------
 * The reference's `E_step(); M_step()` pair (MMM:1341, 1357) without the convergence logic: E-step under
 * params_dev (new labels into assign_dev, switch count), then M-step of the new labels into params_dev.
 * The statistics are accumulated inside the E-step kernel (one pass over the data).  phases: bit 0 = the
 * E-step + statistics + reduction (+ all-reduce) part, bit 1 = the parameter solve; 3 = both (the split
 * only exists so that a benchmark can time the dominant kernel on its own).  Synchronises if info != NULL.
 *
 * Delta statistics (mtm_em_step and the iterations of mtm_em_fit; single fits with hidden states).  The reference
 * recomputes every cluster's sums from all of its members in every M-step (MMM:1254-1308).  Late EM iterations move
 * few trajectories, and the sums of the new labels are those of the old labels plus / minus the moments of the movers.
 * The handle keeps, on the device, the reduced sums (as double-double pairs) of the last two label vectors its fused
 * E-step produced, keyed by a 128-bit hash of the labels.  An iteration whose starting labels hash to a cached entry
 * whose own iteration changed at most n/8 labels only visits the trajectories that change cluster; otherwise it
 * sums everything and refills the cache.  The decision is taken on the device, per rank, without a host round
 * trip; labels changed by the caller simply miss the cache.  A mover's moments are formed by the same operations
 * as in the full pass, so the two modes differ only in the order trajectories are added (as grids and GPU counts
 * do): same labels, parameters equal to rounding.  The cache is dropped by every mtm_set_data (re-bind after
 * changing the resident buffers in place), by a new pivot and at the start of mtm_em_fit.
 * mtm_set_delta_stats(h, mode): 0 = off (every iteration sums everything, no label hashes), 1 = automatic (default:
 * on unless the set is so small — n n_c < 2e6 — that the iteration is bound by launch latency and three more small
 * launches cost more than they save), 2 = on for every size.  Drops the cache. */
int mtm_set_delta_stats(mtm_handle *h, int enable);
int mtm_em_step(mtm_handle *h, int n_c, double alpha, int32_t *assign_dev, double *params_dev, int phases,
                mtm_estep_info *info);

/* ---- one EM iteration, end to end from HOST buffers -----------------------------------------------
 * Same result as mtm_em_step(phases = 3) on data that lives in host memory (pinned for full PCIe rate):
 * the trajectory axis is cut into n_slabs slabs; slab k+1 is uploaded on a private copy stream
 * (2-D copies straight into its place in the (T, n, d|l) device buffers, no repacking) while slab k runs
 * the fused E-step + statistics kernel, and slab k's new labels are downloaded while slab k+1 computes.
 *   z_host, x_host      (T, n, d|l) float64 host arrays;  z_dev, x_dev: device buffers of the same shape
 *   assign_host         (n) int32 in/out (old labels in, new labels out);  assign_dev: (n) int32 device scratch
 *   params_dev          (n_c, param_stride) in/out: parameters of the previous iteration in, new ones out
 *   params_host         (n_c, param_stride) out
 *   info                nullable: n_switches, min_count, status.   Synchronises. */
int mtm_em_step_host(mtm_handle *h, const double *z_host, const double *x_host, int T, int64_t n, int d, int l,
                     double *z_dev, double *x_dev, int n_c, double alpha, int32_t *assign_host,
                     int32_t *assign_dev, double *params_dev, double *params_host, int n_slabs,
                     mtm_estep_info *info);

/* ---- EM fit --------------------------------------------------------------------------------------
 * train (MMM:1310-1366) on the resident set, loop resident on the device:
 *   guard -> M -> { E ; stop if 0 switches ; guard ; M } for at most n_steps E-steps.
 * assign_dev (n) int32 in/out: initial assignment in, final assignment out.
 * params_dev (n_c, param_stride) out: parameters of the last executed M-step (untouched on GUARD0).
 * switches_host (nullable, capacity switches_cap): n_switches after every E-step.
 * With a communicator attached every rank calls this on its shard; one all-reduce of
 * [statistics | n_switches] per iteration; all ranks return identical params/info. */
int mtm_em_fit(mtm_handle *h, int n_c, double alpha, int n_steps, int32_t *assign_dev,
               double *params_dev, mtm_fit_info *info, int64_t *switches_host, int switches_cap);

/* ---- framework_extended: mixtures / classifiers with linear-Gaussian components -----------------------
 * The reference's second mixture class (framework_extended/state_space_model_mixture.py, SSMM) fits every
 * cluster with StateSpaceLinearGaussian.fit (framework_extended/state_space_model_linear_gaussian.py:39-79,
 * SSLG), whose estimates differ from MMLinGaussSS_marginalizable.M_step in two places:
 *   - initial-state mean = np.nanmean per coordinate (every non-NaN entry of z_1 counts, SSLG:43-44); the
 *     covariance still uses the completely finite rows (SSLG:45-47);
 *   - alpha <= 2 eps: util.regress (util/util_state_space.py:33-61) = normal equations with a fixed 1e-6
 *     ridge; alpha > 2 eps: util.regress_alpha (Ridge, :64-94), as in the other class.
 * mtm_mstep_ext: those estimates for all n_c label values in ONE pass over the resident set (the M-step of
 *   SSMM._M_step, SSMM:134-149; the per-class fits of StateSpaceModelClassifier.fit, classifier :33-44; a
 *   single component's fit with n_c = 1).  Same arguments as mtm_mstep.
 * mtm_em_fit_ext: the EM loop of SSMM.fit (SSMM:262-276) resident on the device:
 *     M -> { E ; stop if a cluster got no member (labels of this E-step discarded) ; stop if 0 switches ;
 *            stop if min cluster size <= 3 ; M }
 *   There is no guard before the first M-step (the caller re-draws a near-empty initial assignment, SSMM:251-260).
 *   assign_prev_dev: (n) int32 scratch: the labels before the last executed E-step, which are the model's
 *   labels on MTM_EXIT_EMPTY / MTM_EXIT_ERROR (the reference leaves the assignment untouched when its E-step
 *   raises).  Not sharded: a communicator must not be attached. */
int mtm_mstep_ext(mtm_handle *h, const int32_t *assign_dev, int n_c, double alpha, double *params_dev,
                  mtm_estep_info *info);
int mtm_em_fit_ext(mtm_handle *h, int n_c, double alpha, int n_steps, int32_t *assign_dev, int32_t *assign_prev_dev,
                   double *params_dev, mtm_fit_info *info, int64_t *switches_host, int switches_cap);

/* ---- batched restarts (kernel 4) -------------------------------------------------------------------
 * n_restarts independent EM fits of the resident set with the same n_c, advanced in lock-step: every
 * kernel of an iteration serves all still-running fits at once (one grid row per fit), so a thousand
 * small fits cost a few launches per iteration instead of a few thousand.  This is the loop body of
 * train_with_multiple_random_starts (MMM:1447-1464) and of model-selection.py:45-54 for one n_c; the
 * caller runs it once per n_c of a sweep.
 *   assign_dev   (n_restarts, n) int32 in/out: initial assignments in (e.g. rng.integers for seeds 100+i,
 *                MMM:166-168), final assignments out
 *   params_dev   (n_restarts, n_c, param_stride) out
 *   info_host    [n_restarts]: exit code, E/M-step counts, status of every fit (same rules as mtm_em_fit)
 *   scores_host  (n_restarts, 2) nullable: Q and LL of every final model (MMM:579-666), one launch
 *   score_status_host [n_restarts] nullable: MTM_STATUS_* bits raised while scoring
 *   switches_host (n_restarts, switches_cap) nullable: switch counts per E-step
 * With a communicator attached all ranks call it on their shards. */
int mtm_sweep_fit(mtm_handle *h, int n_c, int n_restarts, double alpha, int n_steps, int32_t *assign_dev,
                  double *params_dev, mtm_fit_info *info_host, double *scores_host, int32_t *score_status_host,
                  int64_t *switches_host, int switches_cap);

/* ---- data-parallel sharding (one process per GPU) ----------------------------------------------
 * The library opens the NCCL that is already loaded in the process (libnccl.so.2).  Rank 0 obtains
 * an id with mtm_comm_unique_id and distributes the 128 bytes by any means (the Python host uses
 * torch.distributed); then every rank calls mtm_comm_init.  Collective calls: mtm_comm_init,
 * mtm_mstep, mtm_em_fit, mtm_score, mtm_comm_destroy. */
int mtm_comm_unique_id(void *id128_out);
int mtm_comm_init(mtm_handle *h, const void *id128, int world_size, int rank);
/* alternative to mtm_comm_init: use a communicator the caller already owns (an ncclComm_t of the same NCCL
 * library, world_size ranks, this process = rank).  The caller keeps ownership; mtm_comm_destroy only detaches. */
int mtm_comm_attach(mtm_handle *h, void *nccl_comm, int world_size, int rank);
int mtm_comm_destroy(mtm_handle *h);
/* Optional, after mtm_comm_init: replace the per-iteration ncclAllReduce of the statistics (6.8 KB at n_c = 8,
 * latency-bound) by a one-shot all-reduce over NVLink peer memory — every rank stores its vector into every
 * peer's mailbox (CUDA IPC mapped), then adds the world vectors in rank order (bit-identical on all ranks).
 *   mtm_p2p_export: allocate this rank's mailbox and return its 64-byte IPC handle;
 *   mtm_p2p_attach: handles_all = world_size x 64 bytes in rank order (exchanged by the caller).
 * Exchanges larger than the mailbox (many batched restarts) keep using NCCL. */
int mtm_p2p_export(mtm_handle *h, void *handle64_out);
int mtm_p2p_attach(mtm_handle *h, const void *handles_all, int world_size, int rank);

/* ---- initial assignment on the device --------------------------------------------------------------
 * Bit-for-bit twin of the reference's random initialisation
 *     np.random.default_rng(seed).integers(low=0, high=n_c, size=n)          [MMM:128, 166-168]
 * pcg_state = {state_hi, state_lo, inc_hi, inc_lo}: the 128-bit PCG64 state and increment of the seeded
 * generator (numpy: np.random.PCG64(seed).state["state"]).  assign_dev (n) int32 out.  Every thread jumps
 * its own copy of the generator ahead; numpy's rare Lemire rejections (each shifts all later elements by
 * one draw) are located first and replayed exactly.  Synchronises. */
int mtm_random_labels(mtm_handle *h, const uint64_t pcg_state[4], int n_c, int64_t n, int32_t *assign_dev,
                      int *n_rejected_host);

/* ---- k-means initial assignment on the device ------------------------------------------------------
 * The reference's init="kmeans" (MMM:132-151): sklearn.cluster.KMeans(n_clusters, init="k-means++",
 * random_state=seed).fit_predict on the FIRST NON-NULL STATE of every trajectory (first step whose d coordinates
 * are all finite, step 0 if there is none; remaining NaN -> column mean).  Same algorithm as sklearn's
 * (column-centred features, k-means++ seeding with 2 + log(k) local trials, Lloyd until no label changes or
 * the squared centre shift is <= rel_tol * mean column variance, final labelling), driven by the SAME random
 * numbers: the caller draws them from numpy's RandomState(seed) exactly as sklearn would —
 *   first_center = rs.choice(n, p = 1/n);   u_host[(c-1) * n_trials + l] = rs.uniform(size = n_trials)[l]  per centre c >= 1
 * — so the device start coincides with sklearn's up to rounding-level ties.
 *   data        resident set (z_dev required); perm_dev: its pattern-order permutation or NULL — seeding follows
 *               the CALLER's trajectory order, like sklearn on the caller's array
 *   assign_dev  (n) int32 out, in the order of `data`
 * Synchronises.  Not sharded. */
typedef struct {
    int32_t n_iter;        /* Lloyd iterations run */
    int32_t converged;     /* 1 = no label changed, 2 = centre shift <= tol, 0 = max_iter reached */
    int32_t empty_cluster; /* a cluster lost all members at some iteration (its centre was kept) */
    int32_t reserved;
    double tol;            /* absolute tolerance used */
} mtm_kmeans_info;
int mtm_kmeans_init(mtm_handle *h, const mtm_data *data, const int64_t *perm_dev, int n_c, int64_t first_center,
                    const double *u_host, int n_trials, int max_iter, double rel_tol, int32_t *assign_dev,
                    mtm_kmeans_info *info);

/* ---- pattern-ordered residency --------------------------------------------------------------------
 * The reference keeps trajectories in the caller's order (MMM:75-81).  For a resident training set the
 * E-step is faster when every aligned group of 32 trajectories shares its per-step pattern of missing
 * hidden-state coordinates: when all of z_t is present the filtered state is a point mass, when exactly one
 * coordinate is missing its covariance has rank one, and a warp that agrees on either skips the generic
 * covariance arithmetic of that step (bit-identical results).
 * mtm_pattern_order: perm_dev (n) int64 out — a stable sort of the trajectories by the key
 *   sum over t < min(T, 8) of s_t 3^(min(T,8)-1-t),  s_t = 2 if all d coordinates of z[t, i, :] are finite,
 *   1 if exactly one is missing, 0 otherwise (step 0 most significant);  perm[p] = index, in the caller's
 *   order, of the trajectory to store at resident position p.  A pure function of the data.  Synchronises.
 * mtm_permute_f64 / _i32: apply it to a (rows, n, w) array (w = 1 for _i32):
 *   inverse == 0:  dst[r, p, :] = src[r, perm[p], :]   (caller's order -> resident order)
 *   inverse != 0:  dst[r, perm[p], :] = src[r, p, :]   (resident order -> caller's order)
 *   e.g. z (T, n, d): rows = T, w = d;  labels (n): rows = 1;  ll / probs (n_c, n): rows = n_c, w = 1.
 *   Out of place only.  Enqueued on the handle's stream.
 * Every other entry point is order-agnostic: it works on whatever order the bound data and the label
 * array share.
 * mtm_pattern_order and mtm_kmeans_init work in a grow-only workspace owned by the handle (about 10 bytes per
 * trajectory for the order, 2 (n, d) copies for k-means): allocating and freeing it per call cost tens to hundreds
 * of milliseconds once peer GPUs have the device mapped.  mtm_release_workspace gives it back (synchronises);
 * mtm_destroy does too. */
int mtm_pattern_order(mtm_handle *h, const mtm_data *data, int64_t *perm_dev);
int mtm_release_workspace(mtm_handle *h);
int mtm_permute_f64(mtm_handle *h, const double *src_dev, double *dst_dev, int64_t rows, int64_t n, int w,
                    const int64_t *perm_dev, int inverse);
int mtm_permute_i32(mtm_handle *h, const int32_t *src_dev, int32_t *dst_dev, int64_t rows, int64_t n,
                    const int64_t *perm_dev, int inverse);

/* ---- cache key of a large data set -----------------------------------------------------------------
 * The reference names its model cache after an md5 over all data bytes (MMM:177-188): ~10 s of host time for
 * config 3's 5 GB.  mtm_fingerprint is the device-side stand-in: a 128-bit key of the data set (element values
 * and their positions in the CALLER's order; perm_dev = the resident order's permutation or NULL), one pass at
 * HBM speed, independent of the storage order and of the reduction order.  data->z_dev may be NULL.
 * key_host: 2 x uint64 out.  Synchronises.  Not cryptographic; the reference's md5 stays available on the host. */
int mtm_fingerprint(mtm_handle *h, const mtm_data *data, const int64_t *perm_dev, uint64_t key_host[2]);

/* ---- measurement ---------------------------------------------------------------------------------
 * DFMA-saturating micro-kernel: FP64 peak of this device in TFLOP/s (FMA = 2 flop), best of `reps`,
 * CUDA-event timed.  Denominator of the FP64 roofline (MEASURED_PEAKS.json has none). */
int mtm_fp64_peak(mtm_handle *h, int reps, double *tflops_out);
/* number of kernels this handle has launched since creation (bench.py's gpu_launches) */
int64_t mtm_launch_count(mtm_handle *h);

#ifdef __cplusplus
}
#endif
#endif /* MTM_B200_H */
