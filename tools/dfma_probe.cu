// Probe: DFMA throughput on sm_100a as a function of where the operands come from.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_probe dfma_probe.cu && ./dfma_probe
#include <cstdio>
#include <cuda_runtime.h>

__constant__ double c_tab[4096];

constexpr int K = 8;  // independent accumulators per thread

// (a) one register source, two constant-bank sources
__global__ void __launch_bounds__(256) k_const(double *out, int iters, double a, double b) {
    double x[K];
    for (int k = 0; k < K; ++k) x[k] = threadIdx.x * 1e-3 + k;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int k = 0; k < K; ++k) x[k] = fma(x[k], a, b);
    double s = 0; for (int k = 0; k < K; ++k) s += x[k];
    if (s == 123.456) out[threadIdx.x] = s;
}
// (b) three distinct register sources per DFMA, no operand shared by neighbours
__global__ void __launch_bounds__(256) k_reg3(double *out, int iters, const double *in) {
    double x[K], y[K], w[K];
    for (int k = 0; k < K; ++k) { x[k] = in[threadIdx.x + k]; y[k] = in[threadIdx.x + 8 + k]; w[k] = in[threadIdx.x + 16 + k]; }
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int k = 0; k < K; ++k) x[k] = fma(y[k], w[(k + u) % K], x[k]);
    double s = 0; for (int k = 0; k < K; ++k) s += x[k];
    if (s == 123.456) out[threadIdx.x] = s;
}
// (c) rank-1 update pattern: one multiplicand shared by consecutive DFMAs (operand reuse cache)
__global__ void __launch_bounds__(256) k_shared_g(double *out, int iters, const double *in) {
    double x[K], w[K], g[8];
    for (int k = 0; k < K; ++k) { x[k] = in[threadIdx.x + k]; w[k] = in[threadIdx.x + 16 + k]; g[k] = in[threadIdx.x + 32 + k]; }
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int k = 0; k < K; ++k) x[k] = fma(g[u], w[k], x[k]);
    double s = 0; for (int k = 0; k < K; ++k) s += x[k];
    if (s == 123.456) out[threadIdx.x] = s;
}
// (d) one source from __constant__ with a run-time (warp-uniform) index
__global__ void __launch_bounds__(256) k_cidx(double *out, int iters, const double *in, int base) {
    double x[K], y[K];
    for (int k = 0; k < K; ++k) { x[k] = in[threadIdx.x + k]; y[k] = in[threadIdx.x + 8 + k]; }
    for (int it = 0; it < iters; ++it) {
        const double *t = c_tab + base + (it & 7) * 64;
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int k = 0; k < K; ++k) x[k] = fma(y[k], t[u * 8 + k], x[k]);
    }
    double s = 0; for (int k = 0; k < K; ++k) s += x[k];
    if (s == 123.456) out[threadIdx.x] = s;
}
// (e) one source from shared memory (warp-uniform address), loaded just before use
__global__ void __launch_bounds__(256) k_smem(double *out, int iters, const double *in, int base) {
    __shared__ double tab[1024];
    for (int k = threadIdx.x; k < 1024; k += 256) tab[k] = in[k];
    __syncthreads();
    double x[K], y[K];
    for (int k = 0; k < K; ++k) { x[k] = in[threadIdx.x + k]; y[k] = in[threadIdx.x + 8 + k]; }
    for (int it = 0; it < iters; ++it) {
        const double *t = tab + base + (it & 7) * 64;
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int k = 0; k < K; ++k) x[k] = fma(y[k], t[u * 8 + k], x[k]);
    }
    double s = 0; for (int k = 0; k < K; ++k) s += x[k];
    if (s == 123.456) out[threadIdx.x] = s;
}

// (f) one source from __constant__[uniform idx], each constant shared by NS consecutive DFMAs
template <int NS>
__global__ void __launch_bounds__(256) k_cshare(double *out, int iters, const double *in, int base) {
    double x[K], y[K];
    for (int k = 0; k < K; ++k) { x[k] = in[threadIdx.x + k]; y[k] = in[threadIdx.x + 8 + k]; }
    for (int it = 0; it < iters; ++it) {
        const double *t = c_tab + base + (it & 7) * 64;
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int k = 0; k < K; ++k) x[k] = fma(y[(k + u) % K], t[u * (K / NS) + k / NS], x[k]);
    }
    double s = 0; for (int k = 0; k < K; ++k) s += x[k];
    if (s == 123.456) out[threadIdx.x] = s;
}
// (g) same sharing pattern but the shared operand comes from shared memory (LDS into a vector register)
template <int NS>
__global__ void __launch_bounds__(256) k_sshare(double *out, int iters, const double *in, int base) {
    __shared__ double tab[1024];
    for (int k = threadIdx.x; k < 1024; k += 256) tab[k] = in[k];
    __syncthreads();
    double x[K], y[K];
    for (int k = 0; k < K; ++k) { x[k] = in[threadIdx.x + k]; y[k] = in[threadIdx.x + 8 + k]; }
    for (int it = 0; it < iters; ++it) {
        const double *t = tab + base + (it & 7) * 64;
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int k = 0; k < K; ++k) x[k] = fma(y[(k + u) % K], t[u * (K / NS) + k / NS], x[k]);
    }
    double s = 0; for (int k = 0; k < K; ++k) s += x[k];
    if (s == 123.456) out[threadIdx.x] = s;
}

template <class F> double bench(const char *name, F launch, int grid, int iters) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    double best = 1e30;
    for (int r = 0; r < 6; ++r) {
        cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (r >= 2 && ms < best) best = ms;
    }
    double tf = 2.0 * grid * 256.0 * K * 8 * iters / (best * 1e-3) * 1e-12;
    printf("%-46s %8.3f ms  %7.2f TFLOP/s  (%s)\n", name, best, tf, cudaGetErrorString(cudaGetLastError()));
    return tf;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int iters = 4096;
    double *out, *in; cudaMalloc(&out, 1 << 20); cudaMalloc(&in, 1 << 20); cudaMemset(in, 0, 1 << 20);
    for (int occ : {2, 8}) {
        const int grid = p.multiProcessorCount * occ;
        printf("--- %d blocks of 256 threads per SM (%d warps/SM)\n", occ, occ * 8);
        bench("a: reg x const x const", [&] { k_const<<<grid, 256>>>(out, iters, 0.999999, 1e-9); }, grid, iters);
        bench("b: 3 distinct register sources", [&] { k_reg3<<<grid, 256>>>(out, iters, in); }, grid, iters);
        bench("c: shared multiplicand (rank-1 pattern)", [&] { k_shared_g<<<grid, 256>>>(out, iters, in); }, grid, iters);
        bench("d: one source from __constant__[uniform idx]", [&] { k_cidx<<<grid, 256>>>(out, iters, in, 0); }, grid, iters);
        bench("e: one source from shared memory (broadcast)", [&] { k_smem<<<grid, 256>>>(out, iters, in, 0); }, grid, iters);
        bench("f4: __constant__ operand shared by 4 DFMAs", [&] { k_cshare<4><<<grid, 256>>>(out, iters, in, 0); }, grid, iters);
        bench("f8: __constant__ operand shared by 8 DFMAs", [&] { k_cshare<8><<<grid, 256>>>(out, iters, in, 0); }, grid, iters);
        bench("g4: shared-memory operand shared by 4 DFMAs", [&] { k_sshare<4><<<grid, 256>>>(out, iters, in, 0); }, grid, iters);
        bench("g8: shared-memory operand shared by 8 DFMAs", [&] { k_sshare<8><<<grid, 256>>>(out, iters, in, 0); }, grid, iters);
    }
    return 0;
}
