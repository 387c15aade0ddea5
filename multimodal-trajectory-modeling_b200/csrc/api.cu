// C ABI of mtm_b200 (see include/mtm_b200.h): handle, scratch, the stream-resident EM loop and
// the NCCL plumbing.  No C++ exception crosses the boundary.
#include <dlfcn.h>

#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <string>

#include "common.cuh"

#ifndef MTM_CONST_DOUBLES
#define MTM_CONST_DOUBLES 7680  // capacity of the constant-bank parameter store (filter.cuh)
#endif

using namespace mtm;

// ---- minimal NCCL surface, resolved at run time from the libnccl.so.2 already in the process ----
extern "C" {
typedef struct ncclComm *ncclComm_t;
typedef struct {
    char internal[128];
} ncclUniqueId;
}
namespace {
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
};
NcclApi g_nccl;
constexpr int kNcclFloat64 = 8, kNcclSum = 0;

bool load_nccl(std::string &err) {
    if (g_nccl.ok) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        g_nccl.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) {
        err = std::string("cannot dlopen libnccl.so.2: ") + dlerror();
        return false;
    }
#define MTM_SYM(field, name)                                                   \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(g_nccl.lib, name)); \
    if (!g_nccl.field) {                                                       \
        err = std::string("NCCL symbol missing: ") + name;                     \
        return false;                                                          \
    }
    MTM_SYM(GetUniqueId, "ncclGetUniqueId")
    MTM_SYM(CommInitRank, "ncclCommInitRank")
    MTM_SYM(AllReduce, "ncclAllReduce")
    MTM_SYM(CommDestroy, "ncclCommDestroy")
    MTM_SYM(GetErrorString, "ncclGetErrorString")
#undef MTM_SYM
    g_nccl.ok = true;
    return true;
}
std::string g_create_error;
}  // namespace

struct mtm_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    int sm_count = 148;
    mtm_data data{};
    bool has_data = false;
    Ctrl *d_ctrl = nullptr;  // [ctrl_cap]: one control block per concurrently running fit
    Ctrl *h_ctrl = nullptr;  // pinned mirror
    int ctrl_cap = 0;
    double *d_partials = nullptr;
    size_t partials_cap = 0;
    double *d_statsbuf = nullptr;
    size_t statsbuf_cap = 0;
    double *d_score = nullptr;  // [2 outputs | partials]
    size_t score_cap = 0;
    unsigned long long *d_counts = nullptr;  // MTM_MAX_CLUSTERS
    unsigned long long *h_counts = nullptr;  // pinned
    long long *d_hist = nullptr;
    int hist_cap = 0;
    double *h_pin = nullptr;  // small pinned scratch (16 doubles)
    double *h_pin32 = nullptr;  // pinned scratch of ensure_pivot (32 doubles)
    cudaStream_t copy_stream = nullptr;  // private upload stream of the host-streaming path
    cudaStream_t down_stream = nullptr;  // private download stream (PCIe is full duplex)
    cudaEvent_t ev_up[64] = {};          // slab k uploaded
    cudaEvent_t ev_done[64] = {};        // slab k computed
    cudaEvent_t ev_free = nullptr;       // staging buffers free again (previous step finished)
    bool stream_ready = false;
    ncclComm_t comm = nullptr;
    bool comm_owned = true;  // created by mtm_comm_init (destroyed with the handle) or attached by the caller
    int world = 1, rank = 0;
    // peer-memory all-reduce (p2p.cu)
    P2PView p2p{};
    bool p2p_on = false;
    double *p2p_mailbox = nullptr;   // own mailbox (cudaMalloc, exported through IPC)
    void *p2p_mapped[kMaxPeers] = {};  // peers' mailboxes opened here
    unsigned int *p2p_counter = nullptr;
    unsigned int p2p_seq = 0;
    // pivot of the resident set: the point the M-step moments are taken about (ensure_pivot)
    double pivot[2 * MTM_MAX_DIM] = {};
    bool pivot_valid = false;
    int pivot_world = 0;
    double *d_pivot = nullptr;  // [32 sums/counts | scratch]
    size_t pivot_cap = 0;
    double *d_nanmean = nullptr;  // framework_extended: [(n_c, 2d) sums/counts | scratch]
    size_t nanmean_cap = 0;
    long long *d_rng = nullptr;  // mtm_random_labels: rejection list [kRngCap] + counter, allocated on first use
    long long *h_rng = nullptr;  // pinned mirror
    // grow-only workspace of mtm_pattern_order / mtm_kmeans_init (cudaMalloc + cudaFree per call cost tens to
    // hundreds of ms once peers have the device mapped); mtm_release_workspace or mtm_destroy frees it
    unsigned char *d_work = nullptr;
    size_t work_cap = 0;
    // delta statistics (common.cuh StatsCache): device-side cache of the reduced sums of the last two label vectors
    StatsCache *d_cache = nullptr;
    double *d_bases = nullptr;  // 2 entries x n_vals x (hi, lo)
    size_t bases_cap = 0;
    int delta_mode = 1;  // mtm_set_delta_stats: 0 off, 1 on where it pays (not for launch-bound sizes), 2 always
    bool cache_in_use = false;  // the last fused iteration went through the statistics cache
    std::string err;
    int64_t launches = 0;
};

namespace {

int fail(mtm_handle *h, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    else g_create_error = buf;
    return code;
}

#define MTM_CUDA(h, expr)                                                                      \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) return fail(h, -2, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(_e), \
                                           __FILE__, __LINE__, #expr);                         \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

template <class T>
int ensure(mtm_handle *h, T *&ptr, size_t &cap, size_t need) {
    if (need <= cap) return 0;
    if (ptr) MTM_CUDA(h, cudaFree(ptr));
    ptr = nullptr;
    cap = 0;
    MTM_CUDA(h, cudaMalloc(reinterpret_cast<void **>(&ptr), need * sizeof(T)));
    cap = need;
    return 0;
}

int check_dims(mtm_handle *h, const mtm_data *dt, int n_c) {
    if (!dt || !dt->x_dev) return fail(h, -1, "no data: bind a set with mtm_set_data or pass one");
    if (dt->d < 1 || dt->d > MTM_MAX_DIM || dt->l < 1 || dt->l > MTM_MAX_DIM)
        return fail(h, -1, "dimensions d=%d l=%d outside [1, %d]", dt->d, dt->l, MTM_MAX_DIM);
    if (dt->T < 1 || dt->n < 1) return fail(h, -1, "empty data set (T=%d, n=%lld)", dt->T, (long long)dt->n);
    if (dt->ld != 0 && dt->ld < dt->n) return fail(h, -1, "ld=%lld smaller than n=%lld", (long long)dt->ld, (long long)dt->n);
    if (n_c < 1 || n_c > MTM_MAX_CLUSTERS) return fail(h, -1, "n_c=%d outside [1, %d]", n_c, MTM_MAX_CLUSTERS);
    if ((dt->d % 2 == 0 && dt->z_dev && (reinterpret_cast<uintptr_t>(dt->z_dev) & 15)) ||
        (dt->l % 2 == 0 && (reinterpret_cast<uintptr_t>(dt->x_dev) & 15)))
        return fail(h, -1, "z_dev / x_dev must be 16-byte aligned");
    return 0;
}

int ensure_ctrl(mtm_handle *h, int R) {
    if (R <= h->ctrl_cap) return 0;
    MTM_CUDA(h, cudaStreamSynchronize(h->stream));
    if (h->d_ctrl) MTM_CUDA(h, cudaFree(h->d_ctrl));
    if (h->h_ctrl) MTM_CUDA(h, cudaFreeHost(h->h_ctrl));
    h->d_ctrl = nullptr;
    h->h_ctrl = nullptr;
    h->ctrl_cap = 0;
    MTM_CUDA(h, cudaMalloc(reinterpret_cast<void **>(&h->d_ctrl), sizeof(Ctrl) * R));
    MTM_CUDA(h, cudaMallocHost(reinterpret_cast<void **>(&h->h_ctrl), sizeof(Ctrl) * R));
    h->ctrl_cap = R;
    return 0;
}

bool can_fuse(const mtm_data *dt, int n_c) {
    // up to 110 KB of shared memory per block (2 blocks/SM) the fused kernel still beats a separate statistics pass
    return dt->z_dev != nullptr && estep_smem(n_c, dt->T, dt->d, dt->l, true) <= 110 * 1024;
}

int enqueue_estep(mtm_handle *h, const mtm_data *dt, const double *params, int n_c, int T0, int32_t *assign,
                  int count_switches, double *ll, double *probs, double *prenorm, unsigned long long *counts,
                  double *score_partials, int check_done, int *grid_out, bool fused = false, int R = 1,
                  StatsCache *cache = nullptr) {
    EStepArgs a{};
    a.z = dt->z_dev;
    a.x = dt->x_dev;
    a.n = dt->n;
    a.ld = dt->ld > 0 ? dt->ld : dt->n;
    a.T = dt->T;
    a.T0 = T0 < 1 ? 1 : (T0 > dt->T ? dt->T : T0);
    a.d = dt->d;
    a.l = dt->l;
    a.params = params;
    a.n_c = n_c;
    a.assign = assign;
    a.count_switches = count_switches;
    a.ll = ll;
    a.probs = probs;
    a.prenorm = prenorm;
    a.ctrl = h->d_ctrl;
    a.counts = counts;
    a.score_partials = score_partials;
    a.check_done = check_done;
    a.R = R;
    int grid = estep_grid(dt->n, h->sm_count, fused);
    if (fused && R > 1 && grid > h->sm_count * 6) grid = h->sm_count * 6;  // R x grid partial blocks: keep it small
    if (grid_out) *grid_out = grid;
    if (fused) {
        const size_t n_vals = (size_t)n_c * stats_stride(dt->d, dt->l);
        if (int rc = ensure(h, h->d_partials, h->partials_cap, 2 * n_vals * ((size_t)grid + kReduceSlices) * R)) return rc;
        a.stats_partials = h->d_partials;
        memcpy(a.pivot, h->pivot, sizeof(a.pivot));  // valid: the callers of the fused mode ran ensure_pivot
        a.cache = (R == 1) ? cache : nullptr;
        a.delta_grid = estep_delta_grid(dt->n, h->sm_count);
    }
    const size_t smem = estep_smem(n_c, a.T0, a.d, a.l, fused);
    if (smem > 200 * 1024) return fail(h, -1, "E-step shared memory %zu B too large (n_c * T0 too big)", smem);
    // Batched restarts on a large data set: launch in chunks of restarts whose parameter rows fit the
    // constant bank together, so that every chunk runs the fast uniform-operand kernel; for small data
    // sets (launch-latency bound) one launch serves all restarts from shared-memory parameter copies.
    const size_t row = (size_t)n_c * param_stride(a.d, a.l);
    int chunk = R;
    if (R > 1 && row * R > MTM_CONST_DOUBLES && row <= MTM_CONST_DOUBLES && (double)dt->n * n_c >= 2e6)
        chunk = (int)(MTM_CONST_DOUBLES / row);
    const size_t nv2 = fused ? 2 * (size_t)n_c * stats_stride(dt->d, dt->l) : 0;
    for (int r0 = 0; r0 < R; r0 += chunk) {
        EStepArgs b = a;
        b.R = (R - r0 < chunk) ? (R - r0) : chunk;
        b.params = a.params + (size_t)r0 * row;
        b.ctrl = a.ctrl + r0;
        if (a.assign) b.assign = a.assign + (size_t)r0 * dt->n;
        if (a.counts) b.counts = a.counts + (size_t)r0 * MTM_MAX_CLUSTERS;
        if (a.score_partials) b.score_partials = a.score_partials + 2 * (size_t)r0 * grid;
        if (a.stats_partials) b.stats_partials = a.stats_partials + (size_t)r0 * grid * nv2;
        MTM_CUDA(h, launch_estep(b, grid, smem, h->stream));
        h->launches += b.cache ? 2 : 1;  // with the cache: estep_kernel and its delta twin, one of which returns at once
    }
    return 0;
}

// forget the cached statistics (new data, new pivot, resized buffers); stream ordered
int invalidate_stats_cache(mtm_handle *h) {
    if (h->d_cache) MTM_CUDA(h, cudaMemsetAsync(h->d_cache, 0, sizeof(StatsCache), h->stream));
    return 0;
}

// the cache of a single fit's fused iterations, or nullptr when delta statistics cannot be used for this shape
int stats_cache_for(mtm_handle *h, int n_c, StatsCache **out) {
    *out = nullptr;
    const mtm_data &dt = h->data;
    if (h->delta_mode == 0) return 0;
    // small sets are launch-latency bound (see the batch size of mtm_em_fit): three more launches per iteration cost
    // more than the statistics pass they could save
    if (h->delta_mode == 1 && (double)dt.n * n_c < 2e6) return 0;
    if (stats_delta_smem_bytes(n_c, dt.d, dt.l) > estep_smem(n_c, dt.T, dt.d, dt.l, true) - estep_smem(n_c, dt.T, dt.d, dt.l, false))
        return 0;  // the per-warp accumulators do not fit the fused kernel's statistics region
    if (!h->d_cache) {
        MTM_CUDA(h, cudaMalloc(reinterpret_cast<void **>(&h->d_cache), sizeof(StatsCache)));
        MTM_CUDA(h, cudaMemsetAsync(h->d_cache, 0, sizeof(StatsCache), h->stream));
    }
    const size_t need = 4 * (size_t)n_c * stats_stride(dt.d, dt.l);
    if (need > h->bases_cap) {
        MTM_CUDA(h, cudaStreamSynchronize(h->stream));
        if (int rc = ensure(h, h->d_bases, h->bases_cap, need)) return rc;
        if (int rc = invalidate_stats_cache(h)) return rc;
    }
    *out = h->d_cache;
    return 0;
}

constexpr int kP2PCap = 65536;  // doubles per rank and parity in a mailbox (512 KB): 32 batched restarts at n_c = 12 (40 832) fit

int allreduce_stats(mtm_handle *h, double *buf, size_t count, int R = 1) {
    if (!h->comm) return 0;
    if (h->p2p_on && count <= (size_t)kP2PCap) {
        h->p2p_seq += 1;
        MTM_CUDA(h, launch_p2p_allreduce(buf, (int)count, h->p2p, h->p2p_seq, h->p2p_counter, h->d_ctrl, R, h->stream));
        h->launches += 2;
        return 0;
    }
    const int r = g_nccl.AllReduce(buf, buf, count, kNcclFloat64, kNcclSum, h->comm, h->stream);
    if (r != 0) return fail(h, -3, "ncclAllReduce failed: %s", g_nccl.GetErrorString(r));
    return 0;
}

// The pivot of the resident set = per-feature mean of its finite entries (over all ranks when a communicator is
// attached), truncated to 12 mantissa bits.  The M-step statistics are accumulated about it (stats_tile.cuh) and
// finalize_kernel solves on the centred moments, so data with |mean| >> spread lose no digits (the reference works
// on the rows themselves, MMM:1194-1199, 1267-1270, and has no such cancellation).  Any common point would give
// the same parameters up to rounding; what matters is that every rank uses the SAME one, hence: computed lazily at
// the first statistics pass after mtm_set_data — a collective call on every rank — and, with a communicator,
// never skipped on local information.  Truncation makes the value insensitive to the summation order (it is the
// same for every GPU count unless a mean sits on a truncation boundary).
int ensure_pivot(mtm_handle *h) {
    const int world = h->comm ? h->world : 1;
    if (h->pivot_valid && h->pivot_world == world) return 0;
    const mtm_data &dt = h->data;
    const size_t need = 32 + pivot_scratch_doubles(h->sm_count);
    if (int rc = ensure(h, h->d_pivot, h->pivot_cap, need)) return rc;
    const int64_t ld = dt.ld > 0 ? dt.ld : dt.n;
    MTM_CUDA(h, launch_pivot_sums(dt.z_dev, dt.n, ld, dt.T, dt.d, h->sm_count, h->d_pivot + 32, h->d_pivot, h->stream));
    MTM_CUDA(h, launch_pivot_sums(dt.x_dev, dt.n, ld, dt.T, dt.l, h->sm_count, h->d_pivot + 32, h->d_pivot + 16, h->stream));
    h->launches += 4;
    if (int rc = allreduce_stats(h, h->d_pivot, 32)) return rc;
    double host[32];
    MTM_CUDA(h, cudaMemcpyAsync(h->h_pin32, h->d_pivot, sizeof(host), cudaMemcpyDeviceToHost, h->stream));
    MTM_CUDA(h, cudaStreamSynchronize(h->stream));
    memcpy(host, h->h_pin32, sizeof(host));
    for (int part = 0; part < 2; ++part) {
        for (int j = 0; j < MTM_MAX_DIM; ++j) {
            const double sum = host[16 * part + j], cnt = host[16 * part + MTM_MAX_DIM + j];
            double v = cnt > 0.0 ? sum / cnt : 0.0;
            uint64_t bits;
            memcpy(&bits, &v, 8);
            bits &= ~((uint64_t(1) << 40) - 1);  // keep sign, exponent and the top 12 mantissa bits
            memcpy(&v, &bits, 8);
            h->pivot[MTM_MAX_DIM * part + j] = (v == v && v - v == 0.0) ? v : 0.0;
        }
    }
    h->pivot_valid = true;
    h->pivot_world = world;
    return invalidate_stats_cache(h);  // cached sums were taken about the old pivot
}

// block partials in h->d_partials (from `grid` blocks) -> h->d_statsbuf [n_c*sstride | n_switches | n_local],
// all-reduced over the communicator
int enqueue_stats_reduce(mtm_handle *h, int n_c, int grid, int check_done, int with_switches, int R = 1,
                         const StatsCache *cache = nullptr) {
    const mtm_data &dt = h->data;
    const int n_vals = n_c * stats_stride(dt.d, dt.l);
    if (int rc = ensure(h, h->d_statsbuf, h->statsbuf_cap, ((size_t)n_vals + kTail) * R)) return rc;
    MTM_CUDA(h, launch_stats_reduce(h->d_partials, grid, n_vals, h->d_statsbuf,
                                    h->d_partials + 2 * (size_t)R * grid * n_vals, h->d_ctrl, check_done,
                                    with_switches, (double)dt.n, R, h->stream, cache, h->d_bases,
                                    estep_delta_grid(dt.n, h->sm_count)));
    h->launches += 2;
    return allreduce_stats(h, h->d_statsbuf, ((size_t)n_vals + kTail) * R, R);
}

// statistics of the resident set under `assign` (stand-alone kernel) + reduce + all-reduce
int enqueue_stats(mtm_handle *h, const int32_t *assign, int n_c, int check_done, int with_switches, int R = 1,
                  bool raw = false) {
    const mtm_data &dt = h->data;
    if (!raw)
        if (int rc = ensure_pivot(h)) return rc;
    const size_t n_vals = (size_t)n_c * stats_stride(dt.d, dt.l);
    const int grid = stats_grid(dt.n, h->sm_count);
    const size_t smem = stats_smem(n_c, dt.d, dt.l);
    if (smem > 220 * 1024)
        return fail(h, -1, "statistics kernel needs %zu B of shared memory (n_c=%d, d=%d, l=%d): too large", smem, n_c,
                    dt.d, dt.l);
    if (int rc = ensure(h, h->d_partials, h->partials_cap, 2 * n_vals * ((size_t)grid + kReduceSlices) * R)) return rc;
    StatsArgs s{};
    s.z = dt.z_dev;
    s.x = dt.x_dev;
    s.n = dt.n;
    s.ld = dt.ld > 0 ? dt.ld : dt.n;
    s.T = dt.T;
    s.d = dt.d;
    s.l = dt.l;
    s.assign = assign;
    s.n_c = n_c;
    s.partials = h->d_partials;
    s.ctrl = h->d_ctrl;
    s.check_done = check_done;
    s.R = R;
    if (!raw) memcpy(s.pivot, h->pivot, sizeof(s.pivot));  // raw: plain moments (pivot 0), mtm_mstats
    MTM_CUDA(h, launch_stats(s, grid, h->stream));
    h->launches += 1;
    return enqueue_stats_reduce(h, n_c, grid, check_done, with_switches, R);
}

// one EM iteration's kernels: E-step (+ fused statistics when possible) -> reduce -> all-reduce
int enqueue_e_and_stats(mtm_handle *h, int n_c, const double *params, int32_t *assign, int check_done, int R = 1) {
    const mtm_data &dt = h->data;
    if (int rc = ensure_pivot(h)) return rc;
    if (can_fuse(&dt, n_c)) {
        // single fit: the statistics cache decides on the device (label hash of the labels this E-step starts from)
        // whether the iteration only visits the trajectories that change cluster
        StatsCache *cache = nullptr;
        if (R == 1)
            if (int rc = stats_cache_for(h, n_c, &cache)) return rc;
        h->cache_in_use = cache != nullptr;
        if (cache) {
            MTM_CUDA(h, launch_label_hash(assign, dt.n, n_c, cache, h->d_ctrl, check_done, 0, 1, h->sm_count, h->stream));
            h->launches += 1;
        }
        int grid = 0;
        if (int rc = enqueue_estep(h, &dt, params, n_c, dt.T, assign, 1, nullptr, nullptr, nullptr, nullptr, nullptr,
                                   check_done, &grid, true, R, cache))
            return rc;
        if (cache) {
            MTM_CUDA(h, launch_label_hash(assign, dt.n, n_c, cache, h->d_ctrl, check_done, 1, 0, h->sm_count, h->stream));
            h->launches += 1;
        }
        return enqueue_stats_reduce(h, n_c, grid, check_done, 1, R, cache);
    }
    if (int rc = enqueue_estep(h, &dt, params, n_c, dt.T, assign, 1, nullptr, nullptr, nullptr, nullptr, nullptr,
                               check_done, nullptr, false, R))
        return rc;
    return enqueue_stats(h, assign, n_c, check_done, 1, R);
}

// framework_extended: np.nanmean statistics of z_1 under `assign` into h->d_nanmean
int enqueue_nanmean(mtm_handle *h, const int32_t *assign, int n_c, int check_done) {
    const mtm_data &dt = h->data;
    const size_t n_out = (size_t)n_c * 2 * dt.d;
    if (int rc = ensure(h, h->d_nanmean, h->nanmean_cap, n_out + nanmean_scratch_doubles(n_c, dt.d, h->sm_count))) return rc;
    MTM_CUDA(h, launch_nanmean(dt.z_dev, dt.n, dt.d, assign, n_c, h->sm_count, h->d_nanmean + n_out, h->d_nanmean,
                               h->d_ctrl, check_done, h->stream));
    h->launches += 2;
    return 0;
}

int enqueue_finalize(mtm_handle *h, int n_c, double alpha, double *params, int mode, int n_steps, int R = 1,
                     int variant = 0) {
    const mtm_data &dt = h->data;
    FinalizeArgs f{};
    f.variant = variant;
    f.nanmean = variant ? h->d_nanmean : nullptr;
    f.stats = h->d_statsbuf;
    f.n_c = n_c;
    f.d = dt.d;
    f.l = dt.l;
    f.n_total = dt.n;
    f.alpha = alpha;
    f.params = params;
    f.ctrl = h->d_ctrl;
    f.mode = mode;
    f.n_steps = n_steps;
    f.switch_hist = h->d_hist;
    f.R = R;
    memcpy(f.pivot, h->pivot, sizeof(f.pivot));
    MTM_CUDA(h, launch_finalize(f, h->d_statsbuf + (size_t)n_c * stats_stride(dt.d, dt.l), h->stream));
    h->launches += 1;
    return 0;
}

int read_ctrl(mtm_handle *h, int R = 1) {
    MTM_CUDA(h, cudaMemcpyAsync(h->h_ctrl, h->d_ctrl, sizeof(Ctrl) * R, cudaMemcpyDeviceToHost, h->stream));
    MTM_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}

}  // namespace

// =================================================================================================
extern "C" {

int mtm_abi_version(void) { return MTM_ABI_VERSION; }

const char *mtm_last_error(mtm_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int mtm_param_stride(int d, int l) { return param_stride(d, l); }
int mtm_stats_stride(int d, int l) { return stats_stride(d, l); }

int mtm_create(int device, void *stream, mtm_handle **out) {
    if (!out) return fail(nullptr, -1, "out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, -2, "no CUDA device available (%s): mtm_b200 has no CPU path",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= count) return fail(nullptr, -1, "device %d out of range [0, %d)", device, count);
    mtm_handle *h = new (std::nothrow) mtm_handle();
    if (!h) return fail(nullptr, -4, "out of host memory");
    h->device = device;
    h->stream = reinterpret_cast<cudaStream_t>(stream);
    DeviceGuard g(device);
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
        delete h;
        return fail(nullptr, -2, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    }
    if (prop.major < 10) {
        delete h;
        return fail(nullptr, -2, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                    prop.major, prop.minor);
    }
    h->sm_count = prop.multiProcessorCount;
    h->ctrl_cap = 1;
    bool ok = cudaMalloc(reinterpret_cast<void **>(&h->d_ctrl), sizeof(Ctrl)) == cudaSuccess &&
              cudaMallocHost(reinterpret_cast<void **>(&h->h_ctrl), sizeof(Ctrl)) == cudaSuccess &&
              cudaMalloc(reinterpret_cast<void **>(&h->d_counts), sizeof(unsigned long long) * MTM_MAX_CLUSTERS) ==
                  cudaSuccess &&
              cudaMallocHost(reinterpret_cast<void **>(&h->h_counts), sizeof(unsigned long long) * MTM_MAX_CLUSTERS) ==
                  cudaSuccess &&
              cudaMallocHost(reinterpret_cast<void **>(&h->h_pin), sizeof(double) * 16) == cudaSuccess &&
              cudaMallocHost(reinterpret_cast<void **>(&h->h_pin32), sizeof(double) * 32) == cudaSuccess &&
              cudaMemset(h->d_ctrl, 0, sizeof(Ctrl)) == cudaSuccess;
    if (!ok) {
        const char *msg = cudaGetErrorString(cudaGetLastError());
        mtm_destroy(h);
        return fail(nullptr, -2, "allocation of control blocks failed: %s", msg);
    }
    *out = h;
    return 0;
}

int mtm_destroy(mtm_handle *h) {
    if (!h) return 0;
    DeviceGuard g(h->device);
    if (h->comm && h->comm_owned && g_nccl.ok) g_nccl.CommDestroy(h->comm);
    if (h->stream_ready) {
        for (int k = 0; k < 64; ++k) {
            cudaEventDestroy(h->ev_up[k]);
            cudaEventDestroy(h->ev_done[k]);
        }
        cudaEventDestroy(h->ev_free);
        cudaStreamDestroy(h->copy_stream);
        cudaStreamDestroy(h->down_stream);
    }
    for (int q = 0; q < kMaxPeers; ++q)
        if (h->p2p_mapped[q]) cudaIpcCloseMemHandle(h->p2p_mapped[q]);
    cudaFree(h->p2p_mailbox);
    cudaFree(h->p2p_counter);
    cudaFree(h->d_rng);
    cudaFreeHost(h->h_rng);
    cudaFree(h->d_work);
    cudaFree(h->d_cache);
    cudaFree(h->d_bases);
    cudaFree(h->d_pivot);
    cudaFree(h->d_nanmean);
    cudaFree(h->d_ctrl);
    cudaFreeHost(h->h_ctrl);
    cudaFree(h->d_partials);
    cudaFree(h->d_statsbuf);
    cudaFree(h->d_score);
    cudaFree(h->d_counts);
    cudaFreeHost(h->h_counts);
    cudaFree(h->d_hist);
    cudaFreeHost(h->h_pin);
    cudaFreeHost(h->h_pin32);
    delete h;
    return 0;
}

int mtm_set_stream(mtm_handle *h, void *stream) {
    if (!h) return -1;
    h->stream = reinterpret_cast<cudaStream_t>(stream);
    return 0;
}

int mtm_set_data(mtm_handle *h, const mtm_data *data) {
    if (!h) return -1;
    if (!data) {
        h->has_data = false;
        return 0;
    }
    if (int rc = check_dims(h, data, 1)) return rc;
    if (!data->z_dev) return fail(h, -1, "the resident training set needs hidden states (z_dev)");
    // the pivot (ensure_pivot) survives a re-bind of the very same buffers, but only without a communicator: with
    // one, every rank must reach the collective of ensure_pivot, so it is never skipped on local information
    const bool same = h->has_data && !h->comm && h->data.z_dev == data->z_dev && h->data.x_dev == data->x_dev &&
                      h->data.n == data->n && h->data.T == data->T && h->data.d == data->d && h->data.l == data->l &&
                      h->data.ld == data->ld;
    if (!same) h->pivot_valid = false;
    h->data = *data;
    h->has_data = true;
    // cached statistics never survive a bind, not even of the same buffers: their contents may have changed, and a
    // label vector that recurs (same seed, new data) must not find the old sums
    DeviceGuard g(h->device);
    return invalidate_stats_cache(h);
}

int mtm_estep(mtm_handle *h, const mtm_data *data, const double *params_dev, int n_c, int T0, int32_t *assign_dev,
              int count_switches, double *ll_dev, double *probs_dev, double *prenorm_dev, mtm_estep_info *info) {
    if (!h) return -1;
    const mtm_data *dt = data ? data : (h->has_data ? &h->data : nullptr);
    if (int rc = check_dims(h, dt, n_c)) return rc;
    if (!params_dev) return fail(h, -1, "params_dev is NULL");
    DeviceGuard g(h->device);
    MTM_CUDA(h, cudaMemsetAsync(h->d_ctrl, 0, sizeof(Ctrl), h->stream));
    MTM_CUDA(h, cudaMemsetAsync(h->d_counts, 0, sizeof(unsigned long long) * MTM_MAX_CLUSTERS, h->stream));
    if (int rc = enqueue_estep(h, dt, params_dev, n_c, T0, assign_dev, count_switches && assign_dev, ll_dev, probs_dev,
                               prenorm_dev, h->d_counts, nullptr, 0, nullptr))
        return rc;
    if (info) {
        MTM_CUDA(h, cudaMemcpyAsync(h->h_counts, h->d_counts, sizeof(unsigned long long) * n_c, cudaMemcpyDeviceToHost,
                                    h->stream));
        if (int rc = read_ctrl(h)) return rc;
        unsigned long long mn = ~0ull;
        for (int c = 0; c < n_c; ++c) mn = h->h_counts[c] < mn ? h->h_counts[c] : mn;
        info->n_switches = (int64_t)h->h_ctrl->n_switches;
        info->min_count = (int64_t)mn;
        info->status = h->h_ctrl->status;
        info->stats_mode = 0;
    }
    return 0;
}

int mtm_estep_prefix(mtm_handle *h, const mtm_data *data, const double *params_dev, int n_c, int T0,
                     double *ll_prefix_dev, int32_t *status_host) {
    if (!h || !ll_prefix_dev) return -1;
    const mtm_data *dt = data ? data : (h->has_data ? &h->data : nullptr);
    if (int rc = check_dims(h, dt, n_c)) return rc;
    if (!params_dev) return fail(h, -1, "params_dev is NULL");
    DeviceGuard g(h->device);
    EStepArgs a{};
    a.z = dt->z_dev;
    a.x = dt->x_dev;
    a.n = dt->n;
    a.ld = dt->ld > 0 ? dt->ld : dt->n;
    a.T = dt->T;
    a.T0 = T0 < 1 ? 1 : (T0 > dt->T ? dt->T : T0);
    a.d = dt->d;
    a.l = dt->l;
    a.params = params_dev;
    a.n_c = n_c;
    a.ctrl = h->d_ctrl;
    const size_t smem = (size_t)n_c * param_stride(a.d, a.l) * 8 + (size_t)n_c * a.T0 * 4 + 16;
    if (smem > 200 * 1024) return fail(h, -1, "shared memory %zu B too large (n_c * T0 too big)", smem);
    MTM_CUDA(h, cudaMemsetAsync(h->d_ctrl, 0, sizeof(Ctrl), h->stream));
    MTM_CUDA(h, launch_estep_prefix(a, ll_prefix_dev, estep_grid(dt->n, h->sm_count, false), h->stream));
    h->launches += 1;
    if (status_host) {
        if (int rc = read_ctrl(h)) return rc;
        *status_host = h->h_ctrl->status;
    }
    return 0;
}

int mtm_score(mtm_handle *h, const mtm_data *data, const double *params_dev, int n_c, double *out_host,
              int32_t *status_host) {
    if (!h || !out_host) return -1;
    const mtm_data *dt = data ? data : (h->has_data ? &h->data : nullptr);
    if (int rc = check_dims(h, dt, n_c)) return rc;
    DeviceGuard g(h->device);
    const int grid = estep_grid(dt->n, h->sm_count, false);
    if (int rc = ensure(h, h->d_score, h->score_cap, (size_t)2 + 2 * grid)) return rc;
    MTM_CUDA(h, cudaMemsetAsync(h->d_ctrl, 0, sizeof(Ctrl), h->stream));
    if (int rc = enqueue_estep(h, dt, params_dev, n_c, dt->T, nullptr, 0, nullptr, nullptr, nullptr, nullptr,
                               h->d_score + 2, 0, nullptr))
        return rc;
    MTM_CUDA(h, launch_score_reduce(h->d_score + 2, grid, h->d_score, 1, h->stream));
    h->launches += 1;
    if (int rc = allreduce_stats(h, h->d_score, 2)) return rc;
    MTM_CUDA(h, cudaMemcpyAsync(h->h_pin, h->d_score, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (int rc = read_ctrl(h)) return rc;
    out_host[0] = h->h_pin[0];
    out_host[1] = h->h_pin[1];
    if (status_host) *status_host = h->h_ctrl->status;
    return 0;
}

int mtm_mstats(mtm_handle *h, const int32_t *assign_dev, int n_c, double *stats_dev) {
    if (!h) return -1;
    if (!h->has_data) return fail(h, -1, "no resident data set (mtm_set_data)");
    if (int rc = check_dims(h, &h->data, n_c)) return rc;
    if (!assign_dev || !stats_dev) return fail(h, -1, "NULL argument");
    DeviceGuard g(h->device);
    // local statistics only: temporarily detach the communicator
    ncclComm_t keep = h->comm;
    h->comm = nullptr;
    const int rc = enqueue_stats(h, assign_dev, n_c, 0, 0, 1, true);  // plain moments (pivot 0)
    h->comm = keep;
    if (rc) return rc;
    const size_t n_vals = (size_t)n_c * stats_stride(h->data.d, h->data.l);
    MTM_CUDA(h, cudaMemcpyAsync(stats_dev, h->d_statsbuf, n_vals * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    return 0;
}

int mtm_finalize(mtm_handle *h, const double *stats_dev, int n_c, int d, int l, int64_t n_total, double alpha,
                 double *params_dev, int32_t *status_dev) {
    if (!h || !stats_dev || !params_dev) return -1;
    if (d < 1 || d > MTM_MAX_DIM || l < 1 || l > MTM_MAX_DIM || n_c < 1 || n_c > MTM_MAX_CLUSTERS)
        return fail(h, -1, "bad dimensions");
    DeviceGuard g(h->device);
    MTM_CUDA(h, cudaMemsetAsync(h->d_ctrl, 0, sizeof(Ctrl), h->stream));
    FinalizeArgs f{};
    f.stats = stats_dev;
    f.n_c = n_c;
    f.d = d;
    f.l = l;
    f.n_total = n_total;
    f.alpha = alpha;
    f.params = params_dev;
    f.ctrl = h->d_ctrl;
    f.mode = 0;
    MTM_CUDA(h, launch_finalize(f, nullptr, h->stream));
    h->launches += 1;
    if (status_dev)
        MTM_CUDA(h, cudaMemcpyAsync(status_dev, &h->d_ctrl->status, sizeof(int32_t), cudaMemcpyDeviceToDevice, h->stream));
    return 0;
}

int mtm_mstep(mtm_handle *h, const int32_t *assign_dev, int n_c, double alpha, double *params_dev,
              mtm_estep_info *info) {
    if (!h) return -1;
    if (!h->has_data) return fail(h, -1, "no resident data set (mtm_set_data)");
    if (int rc = check_dims(h, &h->data, n_c)) return rc;
    if (!assign_dev || !params_dev) return fail(h, -1, "NULL argument");
    DeviceGuard g(h->device);
    MTM_CUDA(h, cudaMemsetAsync(h->d_ctrl, 0, sizeof(Ctrl), h->stream));
    if (int rc = enqueue_stats(h, assign_dev, n_c, 0, 0)) return rc;
    if (int rc = enqueue_finalize(h, n_c, alpha, params_dev, 0, 0)) return rc;
    if (info) {
        if (int rc = read_ctrl(h)) return rc;
        info->n_switches = 0;
        info->min_count = h->h_ctrl->min_count;
        info->status = h->h_ctrl->status;
        info->stats_mode = 0;
    }
    return 0;
}

int mtm_em_step(mtm_handle *h, int n_c, double alpha, int32_t *assign_dev, double *params_dev, int phases,
                mtm_estep_info *info) {
    if (!h) return -1;
    if (!h->has_data) return fail(h, -1, "no resident data set (mtm_set_data)");
    if (int rc = check_dims(h, &h->data, n_c)) return rc;
    if (!assign_dev || !params_dev) return fail(h, -1, "NULL argument");
    DeviceGuard g(h->device);
    if (phases & 1) {
        MTM_CUDA(h, cudaMemsetAsync(h->d_ctrl, 0, sizeof(Ctrl), h->stream));
        if (int rc = enqueue_e_and_stats(h, n_c, params_dev, assign_dev, 0)) return rc;
    }
    if (phases & 2) {
        if (int rc = enqueue_finalize(h, n_c, alpha, params_dev, 0, 0)) return rc;
    }
    if (info) {
        if (int rc = read_ctrl(h)) return rc;
        info->n_switches = (int64_t)h->h_ctrl->n_switches;
        info->min_count = h->h_ctrl->min_count;
        info->status = h->h_ctrl->status;
        info->stats_mode = 0;
        if ((phases & 1) && h->d_cache && h->cache_in_use) {  // how this iteration's statistics were formed (the call has synchronised)
            int mode = 0;
            MTM_CUDA(h, cudaMemcpy(&mode, &h->d_cache->mode, sizeof(int), cudaMemcpyDeviceToHost));
            info->stats_mode = mode;
        }
    }
    return 0;
}

int mtm_set_delta_stats(mtm_handle *h, int enable) {
    if (!h) return -1;
    if (enable < 0 || enable > 2) return fail(h, -1, "mtm_set_delta_stats: 0 (off), 1 (automatic) or 2 (always)");
    h->delta_mode = enable;
    DeviceGuard g(h->device);
    return invalidate_stats_cache(h);
}

int mtm_em_step_host(mtm_handle *h, const double *z_host, const double *x_host, int T, int64_t n, int d, int l,
                     double *z_dev, double *x_dev, int n_c, double alpha, int32_t *assign_host,
                     int32_t *assign_dev, double *params_dev, double *params_host, int n_slabs,
                     mtm_estep_info *info) {
    if (!h) return -1;
    if (!z_host || !x_host || !z_dev || !x_dev || !assign_host || !assign_dev || !params_dev || !params_host)
        return fail(h, -1, "NULL argument");
    mtm_data full{};
    full.z_dev = z_dev;
    full.x_dev = x_dev;
    full.n = n;
    full.T = T;
    full.d = d;
    full.l = l;
    if (int rc = check_dims(h, &full, n_c)) return rc;
    if (!can_fuse(&full, n_c)) return fail(h, -1, "configuration too large for the fused kernel (n_c=%d)", n_c);
    if (n_slabs < 1) n_slabs = 1;
    if (n_slabs > 64) n_slabs = 64;
    if ((int64_t)n_slabs * 128 > n) n_slabs = (int)((n + 127) / 128);
    DeviceGuard g(h->device);
    if (!h->stream_ready) {
        MTM_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        MTM_CUDA(h, cudaStreamCreateWithFlags(&h->down_stream, cudaStreamNonBlocking));
        for (int k = 0; k < 64; ++k) {
            MTM_CUDA(h, cudaEventCreateWithFlags(&h->ev_up[k], cudaEventDisableTiming));
            MTM_CUDA(h, cudaEventCreateWithFlags(&h->ev_done[k], cudaEventDisableTiming));
        }
        MTM_CUDA(h, cudaEventCreateWithFlags(&h->ev_free, cudaEventDisableTiming));
        h->stream_ready = true;
    }
    h->data = full;  // reductions and finalise read the dimensions of the resident set
    h->has_data = true;
    // the data are not on the device yet when the first slab's statistics are taken: plain moments (pivot 0)
    memset(h->pivot, 0, sizeof(h->pivot));
    h->pivot_valid = false;
    const int sstride = stats_stride(d, l);
    const size_t n_vals = (size_t)n_c * sstride;
    // slab boundaries on 128-trajectory tiles (16-byte alignment of every slab start follows)
    const int64_t tiles = (n + 127) / 128;
    int grids[64];
    int64_t start[65];
    int grid_total = 0;
    for (int k = 0; k <= n_slabs; ++k) {
        const int64_t t0 = tiles * k / n_slabs;
        start[k] = t0 * 128 < n ? t0 * 128 : n;
    }
    start[n_slabs] = n;
    for (int k = 0; k < n_slabs; ++k) {
        grids[k] = estep_grid(start[k + 1] - start[k], h->sm_count, true);
        // slab-sized launches: 24 blocks per SM measured better here than the 36 of the resident path (e2e step
        // 95.0 vs 97.6 ms at 10 M trajectories in 16 slabs)
        if (grids[k] > h->sm_count * 24) grids[k] = h->sm_count * 24;
        grid_total += grids[k];
    }
    if (int rc = ensure(h, h->d_partials, h->partials_cap, 2 * n_vals * ((size_t)grid_total + kReduceSlices))) return rc;
    MTM_CUDA(h, cudaMemsetAsync(h->d_ctrl, 0, sizeof(Ctrl), h->stream));
    // the staging buffers may still be read by work enqueued before this call
    MTM_CUDA(h, cudaEventRecord(h->ev_free, h->stream));
    MTM_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->ev_free, 0));
    int goff = 0;
    for (int k = 0; k < n_slabs; ++k) {
        const int64_t s0 = start[k], ns = start[k + 1] - start[k];
        // upload slab k: T rows of ns*d (ns*l) doubles, pitch n*d (n*l) doubles on both sides
        MTM_CUDA(h, cudaMemcpy2DAsync(z_dev + s0 * d, (size_t)n * d * 8, z_host + s0 * d, (size_t)n * d * 8,
                                      (size_t)ns * d * 8, T, cudaMemcpyHostToDevice, h->copy_stream));
        MTM_CUDA(h, cudaMemcpy2DAsync(x_dev + s0 * l, (size_t)n * l * 8, x_host + s0 * l, (size_t)n * l * 8,
                                      (size_t)ns * l * 8, T, cudaMemcpyHostToDevice, h->copy_stream));
        MTM_CUDA(h, cudaMemcpyAsync(assign_dev + s0, assign_host + s0, (size_t)ns * 4, cudaMemcpyHostToDevice,
                                    h->copy_stream));
        MTM_CUDA(h, cudaEventRecord(h->ev_up[k], h->copy_stream));
        MTM_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_up[k], 0));
        EStepArgs a{};
        a.z = z_dev + s0 * d;
        a.x = x_dev + s0 * l;
        a.n = ns;
        a.ld = n;
        a.T = T;
        a.T0 = T;
        a.d = d;
        a.l = l;
        a.params = params_dev;
        a.n_c = n_c;
        a.assign = assign_dev + s0;
        a.count_switches = 1;
        a.ctrl = h->d_ctrl;
        a.R = 1;
        a.stats_partials = h->d_partials + 2 * n_vals * (size_t)goff;
        MTM_CUDA(h, launch_estep(a, grids[k], estep_smem(n_c, T, d, l, true), h->stream));
        h->launches += 1;
        goff += grids[k];
        // download slab k's labels while the next slab computes
        MTM_CUDA(h, cudaEventRecord(h->ev_done[k], h->stream));
        MTM_CUDA(h, cudaStreamWaitEvent(h->down_stream, h->ev_done[k], 0));
        MTM_CUDA(h, cudaMemcpyAsync(assign_host + s0, assign_dev + s0, (size_t)ns * 4, cudaMemcpyDeviceToHost,
                                    h->down_stream));
    }
    if (int rc = enqueue_stats_reduce(h, n_c, grid_total, 0, 1, 1)) return rc;
    if (int rc = enqueue_finalize(h, n_c, alpha, params_dev, 0, 0, 1)) return rc;
    MTM_CUDA(h, cudaMemcpyAsync(params_host, params_dev, (size_t)n_c * param_stride(d, l) * 8, cudaMemcpyDeviceToHost,
                                h->stream));
    if (int rc = read_ctrl(h)) return rc;
    MTM_CUDA(h, cudaStreamSynchronize(h->down_stream));
    if (info) {
        info->n_switches = (int64_t)h->h_ctrl->n_switches;
        info->min_count = h->h_ctrl->min_count;
        info->status = h->h_ctrl->status;
        info->stats_mode = 0;
    }
    return 0;
}

int mtm_em_fit(mtm_handle *h, int n_c, double alpha, int n_steps, int32_t *assign_dev, double *params_dev,
               mtm_fit_info *info, int64_t *switches_host, int switches_cap) {
    if (!h) return -1;
    if (!h->has_data) return fail(h, -1, "no resident data set (mtm_set_data)");
    if (int rc = check_dims(h, &h->data, n_c)) return rc;
    if (!assign_dev || !params_dev || !info) return fail(h, -1, "NULL argument");
    if (n_steps < 0) return fail(h, -1, "n_steps < 0");
    DeviceGuard g(h->device);
    const mtm_data &dt = h->data;
    if (n_steps > h->hist_cap) {
        if (h->d_hist) MTM_CUDA(h, cudaFree(h->d_hist));
        h->d_hist = nullptr;
        h->hist_cap = 0;
        MTM_CUDA(h, cudaMalloc(reinterpret_cast<void **>(&h->d_hist), sizeof(long long) * n_steps));
        h->hist_cap = n_steps;
    }
    MTM_CUDA(h, cudaMemsetAsync(h->d_ctrl, 0, sizeof(Ctrl), h->stream));
    if (int rc = invalidate_stats_cache(h)) return rc;  // a fit starts from nothing it did not compute itself
    // guard + first M-step (MMM:1328-1337)
    if (int rc = enqueue_stats(h, assign_dev, n_c, 0, 0)) return rc;
    if (int rc = enqueue_finalize(h, n_c, alpha, params_dev, 1, n_steps)) return rc;
    // { E ; stop if 0 switches ; guard ; M }: enqueued in batches, the kernels themselves stop
    // working once the device-side `done` flag is up; the host only polls between batches
    const int batch = ((double)dt.n * n_c < 2e6) ? 16 : 2;
    int issued = 0;
    if (int rc = read_ctrl(h)) return rc;
    while (!h->h_ctrl->done && issued < n_steps) {
        const int nb = (n_steps - issued < batch) ? (n_steps - issued) : batch;
        for (int b = 0; b < nb; ++b) {
            if (int rc = enqueue_e_and_stats(h, n_c, params_dev, assign_dev, 1)) return rc;
            if (int rc = enqueue_finalize(h, n_c, alpha, params_dev, 1, n_steps)) return rc;
        }
        issued += nb;
        if (int rc = read_ctrl(h)) return rc;
    }
    const Ctrl &c = *h->h_ctrl;
    info->exit_code = c.exit_code;
    info->n_e_steps = c.n_e_steps;
    info->status = c.status;
    info->n_m_steps = c.n_m_steps;
    info->last_switches = c.last_switches;
    info->min_count = c.min_count;
    if (switches_host && switches_cap > 0 && c.n_e_steps > 0) {
        const int m = c.n_e_steps < switches_cap ? c.n_e_steps : switches_cap;
        static_assert(sizeof(long long) == sizeof(int64_t), "int64 layout");
        MTM_CUDA(h, cudaMemcpyAsync(switches_host, h->d_hist, sizeof(int64_t) * m, cudaMemcpyDeviceToHost, h->stream));
        MTM_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    return 0;
}

int mtm_mstep_ext(mtm_handle *h, const int32_t *assign_dev, int n_c, double alpha, double *params_dev,
                  mtm_estep_info *info) {
    if (!h) return -1;
    if (!h->has_data) return fail(h, -1, "no resident data set (mtm_set_data)");
    if (int rc = check_dims(h, &h->data, n_c)) return rc;
    if (!assign_dev || !params_dev) return fail(h, -1, "NULL argument");
    if (h->comm) return fail(h, -1, "mtm_mstep_ext is not sharded: detach the communicator");
    if (h->data.ld != 0 && h->data.ld != h->data.n) return fail(h, -1, "mtm_mstep_ext needs a contiguous resident set");
    DeviceGuard g(h->device);
    MTM_CUDA(h, cudaMemsetAsync(h->d_ctrl, 0, sizeof(Ctrl), h->stream));
    if (int rc = enqueue_stats(h, assign_dev, n_c, 0, 0)) return rc;
    if (int rc = enqueue_nanmean(h, assign_dev, n_c, 0)) return rc;
    if (int rc = enqueue_finalize(h, n_c, alpha, params_dev, 0, 0, 1, 1)) return rc;
    if (info) {
        if (int rc = read_ctrl(h)) return rc;
        info->n_switches = 0;
        info->min_count = h->h_ctrl->min_count;
        info->status = h->h_ctrl->status;
        info->stats_mode = 0;
    }
    return 0;
}

int mtm_em_fit_ext(mtm_handle *h, int n_c, double alpha, int n_steps, int32_t *assign_dev, int32_t *assign_prev_dev,
                   double *params_dev, mtm_fit_info *info, int64_t *switches_host, int switches_cap) {
    if (!h) return -1;
    if (!h->has_data) return fail(h, -1, "no resident data set (mtm_set_data)");
    if (int rc = check_dims(h, &h->data, n_c)) return rc;
    if (!assign_dev || !assign_prev_dev || !params_dev || !info) return fail(h, -1, "NULL argument");
    if (n_steps < 0) return fail(h, -1, "n_steps < 0");
    if (h->comm) return fail(h, -1, "mtm_em_fit_ext is not sharded: detach the communicator");
    if (h->data.ld != 0 && h->data.ld != h->data.n) return fail(h, -1, "mtm_em_fit_ext needs a contiguous resident set");
    DeviceGuard g(h->device);
    const mtm_data &dt = h->data;
    if (n_steps > h->hist_cap) {
        if (h->d_hist) MTM_CUDA(h, cudaFree(h->d_hist));
        h->d_hist = nullptr;
        h->hist_cap = 0;
        MTM_CUDA(h, cudaMalloc(reinterpret_cast<void **>(&h->d_hist), sizeof(long long) * n_steps));
        h->hist_cap = n_steps;
    }
    MTM_CUDA(h, cudaMemsetAsync(h->d_ctrl, 0, sizeof(Ctrl), h->stream));
    if (int rc = invalidate_stats_cache(h)) return rc;
    MTM_CUDA(h, launch_copy_labels(assign_prev_dev, assign_dev, dt.n, h->d_ctrl, 0, h->stream));
    h->launches += 1;
    // first M-step (SSMM:263): no guard in front of it
    if (int rc = enqueue_stats(h, assign_dev, n_c, 0, 0)) return rc;
    if (int rc = enqueue_nanmean(h, assign_dev, n_c, 0)) return rc;
    if (int rc = enqueue_finalize(h, n_c, alpha, params_dev, 1, n_steps, 1, 1)) return rc;
    const int batch = ((double)dt.n * n_c < 2e6) ? 16 : 2;
    int issued = 0;
    if (int rc = read_ctrl(h)) return rc;
    while (!h->h_ctrl->done && issued < n_steps) {
        const int nb = (n_steps - issued < batch) ? (n_steps - issued) : batch;
        for (int b = 0; b < nb; ++b) {
            // keep the labels this E-step starts from: they stay the model's labels if the step is discarded
            MTM_CUDA(h, launch_copy_labels(assign_prev_dev, assign_dev, dt.n, h->d_ctrl, 1, h->stream));
            h->launches += 1;
            if (int rc = enqueue_e_and_stats(h, n_c, params_dev, assign_dev, 1)) return rc;
            if (int rc = enqueue_nanmean(h, assign_dev, n_c, 1)) return rc;
            if (int rc = enqueue_finalize(h, n_c, alpha, params_dev, 1, n_steps, 1, 1)) return rc;
        }
        issued += nb;
        if (int rc = read_ctrl(h)) return rc;
    }
    const Ctrl &c = *h->h_ctrl;
    info->exit_code = c.exit_code;
    info->n_e_steps = c.n_e_steps;
    info->status = c.status;
    info->n_m_steps = c.n_m_steps;
    info->last_switches = c.last_switches;
    info->min_count = c.min_count;
    if (c.exit_code == MTM_EXIT_EMPTY || c.exit_code == MTM_EXIT_ERROR) {
        // the reference's _E_step raised before it stored anything: the previous labels stand
        MTM_CUDA(h, cudaMemcpyAsync(assign_dev, assign_prev_dev, sizeof(int32_t) * dt.n, cudaMemcpyDeviceToDevice, h->stream));
        MTM_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    if (switches_host && switches_cap > 0 && c.n_e_steps > 0) {
        const int m = c.n_e_steps < switches_cap ? c.n_e_steps : switches_cap;
        MTM_CUDA(h, cudaMemcpyAsync(switches_host, h->d_hist, sizeof(int64_t) * m, cudaMemcpyDeviceToHost, h->stream));
        MTM_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    return 0;
}

int mtm_sweep_fit(mtm_handle *h, int n_c, int n_restarts, double alpha, int n_steps, int32_t *assign_dev,
                  double *params_dev, mtm_fit_info *info_host, double *scores_host, int32_t *score_status_host,
                  int64_t *switches_host, int switches_cap) {
    if (!h) return -1;
    if (!h->has_data) return fail(h, -1, "no resident data set (mtm_set_data)");
    if (int rc = check_dims(h, &h->data, n_c)) return rc;
    if (!assign_dev || !params_dev || !info_host) return fail(h, -1, "NULL argument");
    if (n_restarts < 1 || n_restarts > 65535) return fail(h, -1, "n_restarts=%d outside [1, 65535]", n_restarts);
    if (n_steps < 0) return fail(h, -1, "n_steps < 0");
    DeviceGuard g(h->device);
    const mtm_data &dt = h->data;
    const int R = n_restarts;
    if (int rc = ensure_ctrl(h, R)) return rc;
    const size_t hist_need = (size_t)R * (n_steps > 0 ? n_steps : 1);
    if (hist_need > (size_t)h->hist_cap) {
        if (h->d_hist) MTM_CUDA(h, cudaFree(h->d_hist));
        h->d_hist = nullptr;
        h->hist_cap = 0;
        MTM_CUDA(h, cudaMalloc(reinterpret_cast<void **>(&h->d_hist), sizeof(long long) * hist_need));
        h->hist_cap = (int)hist_need;
    }
    MTM_CUDA(h, cudaMemsetAsync(h->d_ctrl, 0, sizeof(Ctrl) * R, h->stream));
    if (int rc = enqueue_stats(h, assign_dev, n_c, 0, 0, R)) return rc;
    if (int rc = enqueue_finalize(h, n_c, alpha, params_dev, 1, n_steps, R)) return rc;
    const int batch = ((double)dt.n * n_c * R < 2e6) ? 16 : ((double)dt.n * n_c * R < 5e7 ? 4 : 1);
    int issued = 0;
    auto all_done = [&]() {
        for (int r = 0; r < R; ++r)
            if (!h->h_ctrl[r].done) return false;
        return true;
    };
    if (int rc = read_ctrl(h, R)) return rc;
    while (!all_done() && issued < n_steps) {
        const int nb = (n_steps - issued < batch) ? (n_steps - issued) : batch;
        for (int b = 0; b < nb; ++b) {
            if (int rc = enqueue_e_and_stats(h, n_c, params_dev, assign_dev, 1, R)) return rc;
            if (int rc = enqueue_finalize(h, n_c, alpha, params_dev, 1, n_steps, R)) return rc;
        }
        issued += nb;
        if (int rc = read_ctrl(h, R)) return rc;
    }
    for (int r = 0; r < R; ++r) {
        const Ctrl &c = h->h_ctrl[r];
        info_host[r].exit_code = c.exit_code;
        info_host[r].n_e_steps = c.n_e_steps;
        info_host[r].status = c.status;
        info_host[r].n_m_steps = c.n_m_steps;
        info_host[r].last_switches = c.last_switches;
        info_host[r].min_count = c.min_count;
    }
    if (switches_host && switches_cap > 0 && n_steps > 0) {
        for (int r = 0; r < R; ++r) {
            const int m = info_host[r].n_e_steps < switches_cap ? info_host[r].n_e_steps : switches_cap;
            if (m > 0)
                MTM_CUDA(h, cudaMemcpyAsync(switches_host + (size_t)r * switches_cap, h->d_hist + (size_t)r * n_steps,
                                            sizeof(int64_t) * m, cudaMemcpyDeviceToHost, h->stream));
        }
        MTM_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    if (scores_host) {
        // Q and LL of every final model in one launch (e_complete_data_log_lik per restart, MMM:1458-1461)
        const int grid = estep_grid(dt.n, h->sm_count, false);
        if (int rc = ensure(h, h->d_score, h->score_cap, (size_t)2 * R + 2 * (size_t)grid * R)) return rc;
        MTM_CUDA(h, cudaMemsetAsync(h->d_ctrl, 0, sizeof(Ctrl) * R, h->stream));
        if (int rc = enqueue_estep(h, &dt, params_dev, n_c, dt.T, nullptr, 0, nullptr, nullptr, nullptr, nullptr,
                                   h->d_score + 2 * R, 0, nullptr, false, R))
            return rc;
        MTM_CUDA(h, launch_score_reduce(h->d_score + 2 * R, grid, h->d_score, R, h->stream));
        h->launches += 1;
        if (int rc = allreduce_stats(h, h->d_score, (size_t)2 * R, R)) return rc;
        MTM_CUDA(h, cudaMemcpyAsync(scores_host, h->d_score, sizeof(double) * 2 * R, cudaMemcpyDeviceToHost, h->stream));
        if (int rc = read_ctrl(h, R)) return rc;
        if (score_status_host)
            for (int r = 0; r < R; ++r) score_status_host[r] = h->h_ctrl[r].status;
    }
    return 0;
}

int mtm_comm_unique_id(void *id128_out) {
    std::string err;
    if (!id128_out) return -1;
    if (!load_nccl(err)) return fail(nullptr, -3, "%s", err.c_str());
    ncclUniqueId id;
    const int r = g_nccl.GetUniqueId(&id);
    if (r != 0) return fail(nullptr, -3, "ncclGetUniqueId: %s", g_nccl.GetErrorString(r));
    memcpy(id128_out, &id, sizeof(id));
    return 0;
}

int mtm_comm_init(mtm_handle *h, const void *id128, int world_size, int rank) {
    if (!h || !id128) return -1;
    std::string err;
    if (!load_nccl(err)) return fail(h, -3, "%s", err.c_str());
    if (h->comm) return fail(h, -1, "a communicator is already attached");
    DeviceGuard g(h->device);
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    const int r = g_nccl.CommInitRank(&h->comm, world_size, id, rank);
    if (r != 0) {
        h->comm = nullptr;
        return fail(h, -3, "ncclCommInitRank: %s", g_nccl.GetErrorString(r));
    }
    h->comm_owned = true;
    h->world = world_size;
    h->rank = rank;
    h->pivot_valid = false;
    // first collective of a communicator sets up its channels (observed: ~1.7 s): pay it here, not inside a fit
    if (int rc = ensure(h, h->d_score, h->score_cap, 8)) return rc;
    MTM_CUDA(h, cudaMemsetAsync(h->d_score, 0, 8 * sizeof(double), h->stream));
    const int w = g_nccl.AllReduce(h->d_score, h->d_score, 8, kNcclFloat64, kNcclSum, h->comm, h->stream);
    if (w != 0) return fail(h, -3, "ncclAllReduce (warm-up) failed: %s", g_nccl.GetErrorString(w));
    MTM_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int mtm_comm_attach(mtm_handle *h, void *nccl_comm, int world_size, int rank) {
    if (!h || !nccl_comm) return -1;
    if (world_size < 1 || rank < 0 || rank >= world_size) return fail(h, -1, "bad world_size / rank");
    std::string err;
    if (!load_nccl(err)) return fail(h, -3, "%s", err.c_str());
    if (h->comm) return fail(h, -1, "a communicator is already attached");
    h->comm = static_cast<ncclComm_t>(nccl_comm);
    h->comm_owned = false;  // the caller keeps ownership: mtm_comm_destroy only detaches it
    h->pivot_valid = false;
    h->world = world_size;
    h->rank = rank;
    return 0;
}

int mtm_p2p_export(mtm_handle *h, void *handle64_out) {
    if (!h || !handle64_out) return -1;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    DeviceGuard g(h->device);
    if (!h->p2p_mailbox) {
        const size_t bytes = (size_t)2 * kMaxPeers * kP2PCap * sizeof(double) + 2 * kMaxPeers * sizeof(unsigned int);
        MTM_CUDA(h, cudaMalloc(reinterpret_cast<void **>(&h->p2p_mailbox), bytes));
        MTM_CUDA(h, cudaMemset(h->p2p_mailbox, 0, bytes));
        MTM_CUDA(h, cudaMalloc(reinterpret_cast<void **>(&h->p2p_counter), sizeof(unsigned int)));
        MTM_CUDA(h, cudaMemset(h->p2p_counter, 0, sizeof(unsigned int)));
    }
    cudaIpcMemHandle_t ipc;
    MTM_CUDA(h, cudaIpcGetMemHandle(&ipc, h->p2p_mailbox));
    memcpy(handle64_out, &ipc, sizeof(ipc));
    return 0;
}

int mtm_p2p_attach(mtm_handle *h, const void *handles_all, int world_size, int rank) {
    if (!h || !handles_all) return -1;
    if (!h->p2p_mailbox) return fail(h, -1, "call mtm_p2p_export first");
    if (world_size < 2 || world_size > kMaxPeers) return fail(h, -1, "world size %d outside [2, %d]", world_size, kMaxPeers);
    DeviceGuard g(h->device);
    const size_t data_doubles = (size_t)2 * kMaxPeers * kP2PCap;
    for (int q = 0; q < world_size; ++q) {
        void *base = nullptr;
        if (q == rank) {
            base = h->p2p_mailbox;
        } else {
            cudaIpcMemHandle_t ipc;
            memcpy(&ipc, static_cast<const char *>(handles_all) + (size_t)q * 64, sizeof(ipc));
            MTM_CUDA(h, cudaIpcOpenMemHandle(&base, ipc, cudaIpcMemLazyEnablePeerAccess));
            h->p2p_mapped[q] = base;
        }
        h->p2p.data[q] = static_cast<double *>(base);
        h->p2p.flags[q] = reinterpret_cast<unsigned int *>(static_cast<double *>(base) + data_doubles);
    }
    h->p2p.world = world_size;
    h->p2p.rank = rank;
    h->p2p.cap = kP2PCap;
    double timeout_s = 60.0;
    if (const char *e = getenv("MTM_B200_P2P_TIMEOUT_S")) {
        const double v = atof(e);
        if (v > 0.0) timeout_s = v;
    }
    h->p2p.timeout_ns = (unsigned long long)(timeout_s * 1e9);
    // p2p_seq is NOT reset: the mailbox (and the sequence numbers left in its flags) outlives a detach, so the
    // counter stays monotonic per handle.  All ranks attach, exchange and detach together (collective calls), hence
    // their counters agree; a restart from 1 would match the stale flags of an earlier session at once.
    h->p2p_on = true;
    return 0;
}

int mtm_comm_destroy(mtm_handle *h) {
    if (!h) return -1;
    if (h->p2p_on) {
        DeviceGuard g(h->device);
        cudaStreamSynchronize(h->stream);
        for (int q = 0; q < kMaxPeers; ++q)
            if (h->p2p_mapped[q]) {
                cudaIpcCloseMemHandle(h->p2p_mapped[q]);
                h->p2p_mapped[q] = nullptr;
            }
        h->p2p_on = false;
    }
    if (h->comm && g_nccl.ok) {
        DeviceGuard g(h->device);
        cudaStreamSynchronize(h->stream);
        if (h->comm_owned) g_nccl.CommDestroy(h->comm);
    }
    h->comm_owned = true;
    h->comm = nullptr;
    h->world = 1;
    h->rank = 0;
    h->pivot_valid = false;
    return 0;
}

int mtm_random_labels(mtm_handle *h, const uint64_t pcg_state[4], int n_c, int64_t n, int32_t *assign_dev,
                      int *n_rejected_host) {
    if (!h || !pcg_state || !assign_dev) return -1;
    if (n_c < 1 || n < 1) return fail(h, -1, "bad arguments");
    DeviceGuard g(h->device);
    constexpr int kCap = 4096;
    // scratch lives with the handle (one call per restart at scale); every copy is stream-ordered with the kernels
    if (!h->d_rng) {
        MTM_CUDA(h, cudaMalloc(reinterpret_cast<void **>(&h->d_rng), sizeof(long long) * (kCap + 1)));
        MTM_CUDA(h, cudaMallocHost(reinterpret_cast<void **>(&h->h_rng), sizeof(long long) * (kCap + 1)));
    }
    long long *d_list = h->d_rng, *h_list = h->h_rng;
    int *d_cnt = reinterpret_cast<int *>(d_list + kCap);
    int *h_cnt = reinterpret_cast<int *>(h_list + kCap);
    unsigned long long st[4] = {pcg_state[0], pcg_state[1], pcg_state[2], pcg_state[3]};
    int n_rej = 0;
    if (n_c > 1) {
        // The first n + slack draws cover the n accepted ones as long as fewer than `slack` of them are rejected;
        // widen until that holds (rejections are ~n * n_c / 2^32).
        long long slack = 64;
        for (int attempt = 0; attempt < 6; ++attempt) {
            MTM_CUDA(h, cudaMemsetAsync(d_cnt, 0, sizeof(int), h->stream));
            MTM_CUDA(h, launch_rng_reject(st, (unsigned int)n_c, n + slack, d_list, kCap, d_cnt, h->stream));
            h->launches += 1;
            MTM_CUDA(h, cudaMemcpyAsync(h_cnt, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            MTM_CUDA(h, cudaStreamSynchronize(h->stream));
            n_rej = *h_cnt;
            if (n_rej > kCap) return fail(h, -1, "more than %d rejected draws: n too large for this n_c", kCap);
            if (n_rej < slack) break;
            slack = 2 * (long long)n_rej + 64;
        }
        if (n_rej > 0) {
            MTM_CUDA(h, cudaMemcpyAsync(h_list, d_list, sizeof(long long) * n_rej, cudaMemcpyDeviceToHost, h->stream));
            MTM_CUDA(h, cudaStreamSynchronize(h->stream));
            for (int a = 1; a < n_rej; ++a) {  // insertion sort: a handful of entries
                const long long v = h_list[a];
                int b = a - 1;
                while (b >= 0 && h_list[b] > v) {
                    h_list[b + 1] = h_list[b];
                    --b;
                }
                h_list[b + 1] = v;
            }
            MTM_CUDA(h, cudaMemcpyAsync(d_list, h_list, sizeof(long long) * n_rej, cudaMemcpyHostToDevice, h->stream));
        }
    }
    MTM_CUDA(h, launch_rng_emit(st, (unsigned int)n_c, n, d_list, n_rej, assign_dev, h->stream));
    h->launches += 1;
    // the pinned list may be rewritten by the next call: wait for the upload (and the kernel) to finish
    MTM_CUDA(h, cudaStreamSynchronize(h->stream));
    if (n_rejected_host) *n_rejected_host = n_rej;
    return 0;
}

int mtm_pattern_order(mtm_handle *h, const mtm_data *data, int64_t *perm_dev) {
    if (!h || !data || !perm_dev) return -1;
    if (!data->z_dev) return fail(h, -1, "mtm_pattern_order needs hidden states (z_dev)");
    if (data->n < 1 || data->T < 1 || data->d < 1 || data->d > MTM_MAX_DIM) return fail(h, -1, "bad data shape");
    DeviceGuard g(h->device);
    if (int rc = ensure(h, h->d_work, h->work_cap, order_scratch_bytes(data->n, h->sm_count))) return rc;
    void *scratch = h->d_work;
    MTM_CUDA(h, launch_pattern_order(data->z_dev, data->n, data->ld > 0 ? data->ld : data->n, data->T, data->d,
                                     h->sm_count, scratch, perm_dev, h->stream));
    h->launches += (data->T > 4) ? 7 : 4;
    MTM_CUDA(h, cudaStreamSynchronize(h->stream));  // the workspace may be reused from another stream next
    return 0;
}

int mtm_release_workspace(mtm_handle *h) {
    if (!h) return -1;
    DeviceGuard g(h->device);
    MTM_CUDA(h, cudaStreamSynchronize(h->stream));
    if (h->d_work) MTM_CUDA(h, cudaFree(h->d_work));
    h->d_work = nullptr;
    h->work_cap = 0;
    return 0;
}

int mtm_permute_f64(mtm_handle *h, const double *src_dev, double *dst_dev, int64_t rows, int64_t n, int w,
                    const int64_t *perm_dev, int inverse) {
    if (!h || !src_dev || !dst_dev || !perm_dev) return -1;
    if (rows < 0 || n < 0 || w < 1) return fail(h, -1, "bad arguments");
    if (src_dev == dst_dev) return fail(h, -1, "mtm_permute_f64 does not work in place");
    if (rows == 0 || n == 0) return 0;
    DeviceGuard g(h->device);
    MTM_CUDA(h, launch_permute_f64(src_dev, dst_dev, rows, n, w, perm_dev, inverse, h->stream));
    h->launches += 1;
    return 0;
}

int mtm_permute_i32(mtm_handle *h, const int32_t *src_dev, int32_t *dst_dev, int64_t rows, int64_t n,
                    const int64_t *perm_dev, int inverse) {
    if (!h || !src_dev || !dst_dev || !perm_dev) return -1;
    if (rows < 0 || n < 0) return fail(h, -1, "bad arguments");
    if (src_dev == dst_dev) return fail(h, -1, "mtm_permute_i32 does not work in place");
    if (rows == 0 || n == 0) return 0;
    DeviceGuard g(h->device);
    MTM_CUDA(h, launch_permute_i32(src_dev, dst_dev, rows, n, perm_dev, inverse, h->stream));
    h->launches += 1;
    return 0;
}

int mtm_kmeans_init(mtm_handle *h, const mtm_data *data, const int64_t *perm_dev, int n_c, int64_t first_center,
                    const double *u_host, int n_trials, int max_iter, double rel_tol, int32_t *assign_dev,
                    mtm_kmeans_info *info) {
    if (!h || !data || !assign_dev || !info) return -1;
    if (!data->z_dev) return fail(h, -1, "mtm_kmeans_init needs hidden states (z_dev)");
    if (data->n < 1 || data->T < 1 || data->d < 1 || data->d > MTM_MAX_DIM) return fail(h, -1, "bad data shape");
    if (n_c < 1 || n_c > MTM_MAX_CLUSTERS || n_c > data->n) return fail(h, -1, "n_c=%d out of range", n_c);
    if (first_center < 0 || first_center >= data->n) return fail(h, -1, "first_center out of range");
    if (n_c > 1 && (!u_host || n_trials < 1)) return fail(h, -1, "k-means++ needs the uniform draws");
    if (h->comm) return fail(h, -1, "mtm_kmeans_init is not sharded: detach the communicator");
    DeviceGuard g(h->device);
    const int64_t n = data->n;
    const int w = data->d;
    const size_t rows_d = (size_t)n * w;
    const size_t total = 2 * rows_d + kmeans_scratch_doubles(n, w, n_c, h->sm_count) + (size_t)n_c * w + (n + 1) / 2 + 8;
    if (int rc = ensure(h, h->d_work, h->work_cap, total * sizeof(double))) return rc;
    double *buf = reinterpret_cast<double *>(h->d_work);
    double *rows_res = buf, *rows = buf + rows_d, *centers = rows + rows_d, *scratch = centers + (size_t)n_c * w;
    int32_t *labels = reinterpret_cast<int32_t *>(scratch + kmeans_scratch_doubles(n, w, n_c, h->sm_count));
    MTM_CUDA(h, launch_first_rows(data->z_dev, n, data->ld > 0 ? data->ld : n, data->T, w, rows_res, h->sm_count, h->stream));
    h->launches += 1;
    if (perm_dev) {  // seeding works on the caller's order
        MTM_CUDA(h, launch_permute_f64(rows_res, rows, 1, n, w, perm_dev, 1, h->stream));
        h->launches += 1;
    } else {
        MTM_CUDA(h, cudaMemcpyAsync(rows, rows_res, rows_d * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    }
    int n_iter = 0, how = 0, empty = 0;
    double tol = 0.0;
    MTM_CUDA(h, run_kmeans(rows, n, w, n_c, first_center, u_host, n_trials, max_iter, rel_tol, labels, centers, scratch,
                           h->sm_count, h->stream, &n_iter, &how, &empty, &tol, &h->launches));
    if (perm_dev) {
        MTM_CUDA(h, launch_permute_i32(labels, assign_dev, 1, n, perm_dev, 0, h->stream));
        h->launches += 1;
    } else {
        MTM_CUDA(h, cudaMemcpyAsync(assign_dev, labels, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToDevice, h->stream));
    }
    MTM_CUDA(h, cudaStreamSynchronize(h->stream));
    info->n_iter = n_iter;
    info->converged = how;
    info->empty_cluster = empty;
    info->reserved = 0;
    info->tol = tol;
    return 0;
}

int mtm_fingerprint(mtm_handle *h, const mtm_data *data, const int64_t *perm_dev, uint64_t key_host[2]) {
    if (!h || !data || !key_host) return -1;
    if (!data->x_dev || data->n < 1 || data->T < 1) return fail(h, -1, "bad data set");
    if (data->ld != 0 && data->ld != data->n) return fail(h, -1, "mtm_fingerprint needs a contiguous data set");
    DeviceGuard g(h->device);
    if (int rc = ensure(h, h->d_score, h->score_cap, 8)) return rc;
    unsigned long long *acc = reinterpret_cast<unsigned long long *>(h->d_score);
    MTM_CUDA(h, cudaMemsetAsync(acc, 0, 2 * sizeof(unsigned long long), h->stream));
    if (data->z_dev)
        MTM_CUDA(h, launch_fingerprint(data->z_dev, data->T, data->n, data->d, perm_dev, 0x5a5a5a5aull, acc, h->sm_count,
                                       h->stream));
    MTM_CUDA(h, launch_fingerprint(data->x_dev, data->T, data->n, data->l, perm_dev, 0xc3c3c3c3ull, acc, h->sm_count,
                                   h->stream));
    h->launches += data->z_dev ? 2 : 1;
    MTM_CUDA(h, cudaMemcpyAsync(h->h_pin, acc, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    MTM_CUDA(h, cudaStreamSynchronize(h->stream));
    memcpy(key_host, h->h_pin, 2 * sizeof(uint64_t));
    return 0;
}

int mtm_fp64_peak(mtm_handle *h, int reps, double *tflops_out) {
    if (!h || !tflops_out) return -1;
    DeviceGuard g(h->device);
    MTM_CUDA(h, run_fp64_peak(reps < 1 ? 1 : reps, h->sm_count, h->stream, tflops_out));
    h->launches += 2 * ((reps < 1 ? 1 : reps) + 2);
    return 0;
}

int64_t mtm_launch_count(mtm_handle *h) { return h ? h->launches : 0; }

}  // extern "C"
