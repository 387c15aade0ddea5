// All T0 prefix log-likelihoods of every (trajectory, cluster) pair in ONE pass.
//
// The reference's cluster_propensities_over_time / observed_cluster_propensities_over_time
// (framework/marginalizable_mixture_model.py:532-577, 1077-1123) call
// conditional_log_likelihoods_first_T0_steps for T0 = 1..T and every cluster: T (T + 1) / 2 time steps of
// filtering per pair.  The sequential filter of filter.cuh passes through every prefix anyway; this kernel
// writes ll[t, c, i] after each step: T steps per pair, one launch.  Not the hot path (a reporting call): the
// parameters are read from shared memory and there is no statistics fusion, so nothing here constrains the code
// shape of estep.cu.
#include "common.cuh"
#include "filter.cuh"

namespace mtm {

template <int D_, int L_, bool HAS_Z>
__global__ void __launch_bounds__(kEThreads) estep_prefix_kernel(const EStepArgs a, double *__restrict__ pre) {
    using DM = Dims<D_, L_>;
    const DM dm{a.d, a.l};
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int stride = dm.stride();
    double *s_par = reinterpret_cast<double *>(smem_raw);                        // n_c * stride
    uint32_t *s_fin = reinterpret_cast<uint32_t *>(s_par + a.n_c * stride);      // n_c * T0
    const int tid = threadIdx.x;
    for (int k = tid; k < a.n_c * stride; k += kEThreads) s_par[k] = a.params[k];
    __syncthreads();
    for (int c = tid; c < a.n_c; c += kEThreads) mean_finite_masks(dm, SmemSrc{s_par + c * stride}, a.T0, s_fin + c * a.T0);
    __syncthreads();
    const int64_t n = a.n;
    const int64_t n_tiles = (n + kEThreads - 1) / kEThreads;
    const int64_t zs = a.ld * dm.d(), xs = a.ld * dm.l();
    int bad = 0;
    double sink = 0.0;  // valid lanes write into pre, the clamped tail lanes into this
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t i_raw = tile * kEThreads + tid;
        const bool valid = i_raw < n;
        const int64_t i = valid ? i_raw : n - 1;
        const double *z_i = HAS_Z ? a.z + i * dm.d() : nullptr;
        const double *x_i = a.x + i * dm.l();
        for (int c = 0; c < a.n_c; ++c) {
            double *out = valid ? pre + (int64_t)c * n + i : &sink;
            filter_eval<DM, HAS_Z, SmemSrc, true>(dm, z_i, zs, x_i, xs, a.T0, SmemSrc{s_par + c * stride},
                                                  s_fin + c * a.T0, bad, a.tshift, out,
                                                  valid ? (int64_t)a.n_c * n : 0);
        }
    }
    if (bad) atomicOr(&a.ctrl->status, bad);
}

template <int D_, int L_>
static cudaError_t launch_prefix_dl(const EStepArgs &a, double *pre, int grid, cudaStream_t s) {
    const size_t smem = (size_t)a.n_c * param_stride(a.d, a.l) * 8 + (size_t)a.n_c * a.T0 * 4 + 16;
    auto launch = [&](auto k) {
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        k<<<grid, kEThreads, smem, s>>>(a, pre);
        return cudaGetLastError();
    };
    return a.z ? launch(estep_prefix_kernel<D_, L_, true>) : launch(estep_prefix_kernel<D_, L_, false>);
}

// pre: (T0, n_c, n) float64 out
cudaError_t launch_estep_prefix(const EStepArgs &a, double *pre, int grid, cudaStream_t s) {
#define MTM_CASE(DD, LL) \
    if (a.d == DD && a.l == LL) return launch_prefix_dl<DD, LL>(a, pre, grid, s);
    MTM_CASE(4, 4)
    MTM_CASE(2, 4)
    MTM_CASE(2, 3)
    MTM_CASE(1, 1)
    MTM_CASE(3, 3)
    MTM_CASE(2, 2)
    MTM_CASE(5, 3)
#undef MTM_CASE
    return launch_prefix_dl<0, 0>(a, pre, grid, s);
}

}  // namespace mtm
