// Probe: how much FP64 (DFMA) throughput survives when other instructions share the issue slots and the register
// file, at the occupancy of the E-step kernel (3 warps per scheduler) and at higher occupancy.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mix_probe mix_probe.cu && ./mix_probe
// Stream per thread: K independent DFMA chains (x = x * a + b, constant-bank operands: the fastest form) with
// NI integer instructions of kind KIND inserted after every ND DFMAs.
//   KIND 0: IMAD on independent integer registers      KIND 1: MOV-like copies between 64-bit registers (IMAD.MOV pairs)
//   KIND 2: LOP3/ISETP-like predicate work             KIND 3: predicated-OFF DFMAs (issue slot, no arithmetic)
//   KIND 4: 64-bit zeroing (CS2R)   KIND 5: 64-bit select (FSEL pair)   KIND 6: 64-bit immediate move
#include <cstdio>
#include <cuda_runtime.h>

constexpr int K = 8;

// FORM 0: x = x * a + b   (one vector-register read per DFMA: a, b constants / reuse cache)
// FORM 1: x = y * a + x   (two fresh vector-register reads per DFMA: the shape of the E-step's parameter DFMAs,
//                          DFMA R, R, UR, R)
template <int KIND, int ND, int NI, int FORM = 0>
__global__ void __launch_bounds__(128) k_mix(double *out, int iters, double a, double b, int zero) {
    double x[K], y[K];
    for (int k = 0; k < K; ++k) y[k] = 1.0 + threadIdx.x * 1e-6 + k * 1e-3;
    unsigned int q[8];
    double m[4];
    for (int k = 0; k < K; ++k) x[k] = threadIdx.x * 1e-3 + k;
    for (int k = 0; k < 8; ++k) q[k] = threadIdx.x + k;
    for (int k = 0; k < 4; ++k) m[k] = threadIdx.x + 0.5 * k;
    const bool off = zero != 0;  // false at run time, unknown to the compiler
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if (FORM == 0) x[k] = fma(x[k], a, b);
                else x[k] = fma(y[(k + u) % K], a, x[k]);
                if ((u * K + k) % ND == ND - 1) {
#pragma unroll
                    for (int j = 0; j < NI; ++j) {
                        const int r = (u * K + k + j) & 7;
                        if (KIND == 0) q[r] = q[r] * 3u + q[(r + 1) & 7];
                        if (KIND == 1) asm volatile("mov.f64 %0, %1;" : "=d"(m[r & 3]) : "d"(m[(r + 1) & 3]));
                        if (KIND == 2) q[r] = (q[r] & 0x7ff00000u) == 0x7ff00000u ? q[(r + 1) & 7] : (q[r] ^ 0x55u);
                        if (KIND == 3) if (off) x[r] = fma(x[r], b, a);
                        if (KIND == 4) asm volatile("mov.b64 %0, 0;" : "=d"(m[r & 3]));                      // CS2R R, SRZ
                        if (KIND == 5) m[r & 3] = (q[r] & 1u) ? m[(r + 1) & 3] : m[(r + 2) & 3];             // FSEL pair
                        if (KIND == 6) asm volatile("mov.b64 %0, 0x3ff0000000000000;" : "=d"(m[r & 3]));     // immediate moves
                    }
                }
            }
        }
    }
    double s = 0;
    for (int k = 0; k < K; ++k) s += x[k] + y[k];
    unsigned int t = 0;
    for (int k = 0; k < 8; ++k) t += q[k];
    for (int k = 0; k < 4; ++k) s += m[k];
    if (s == 123.456 || t == 0x12345u) out[threadIdx.x] = s + t;
}

template <class F> void bench(const char *name, F launch, int grid, int iters) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    double best = 1e30;
    for (int r = 0; r < 6; ++r) {
        cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (r >= 2 && ms < best) best = ms;
    }
    const double tf = 2.0 * grid * 128.0 * K * 8 * iters / (best * 1e-3) * 1e-12;
    printf("%-58s %8.3f ms  %7.2f TFLOP/s  (%s)\n", name, best, tf, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int iters = 4096;
    double *out; cudaMalloc(&out, 1 << 20);
    for (int occ : {3, 4, 8}) {  // blocks of 128 threads per SM = warps per scheduler
        const int grid = p.multiProcessorCount * occ;
        printf("--- %d warps per scheduler\n", occ);
#define RUN(KIND, ND, NI, label) bench(label, [&] { k_mix<KIND, ND, NI><<<grid, 128>>>(out, iters, 0.999999, 1e-9, 0); }, grid, iters)
        RUN(0, 64, 0, "DFMA only");
        RUN(0, 4, 1, "IMAD: 1 per 4 DFMA");
        RUN(0, 2, 1, "IMAD: 1 per 2 DFMA");
        RUN(0, 1, 1, "IMAD: 1 per DFMA");
        RUN(0, 1, 2, "IMAD: 2 per DFMA");
        RUN(1, 2, 1, "64-bit register copy: 1 per 2 DFMA");
        RUN(1, 1, 1, "64-bit register copy: 1 per DFMA");
        RUN(2, 1, 1, "LOP3/ISETP/SEL: 1 group per DFMA");
        RUN(3, 2, 1, "predicated-off DFMA: 1 per 2 DFMA");
        RUN(3, 1, 1, "predicated-off DFMA: 1 per DFMA");
#define RUN2(KIND, ND, NI, label) bench(label, [&] { k_mix<KIND, ND, NI, 1><<<grid, 128>>>(out, iters, 0.999999, 1e-9, 0); }, grid, iters)
        RUN2(0, 64, 0, "2-read DFMA only");
        RUN2(2, 2, 1, "2-read DFMA + LOP3/ISETP/SEL group: 1 per 2 DFMA");
        RUN2(2, 1, 1, "2-read DFMA + LOP3/ISETP/SEL group: 1 per DFMA");
        RUN2(0, 4, 1, "2-read DFMA + IMAD: 1 per 4 DFMA");
        RUN2(0, 2, 1, "2-read DFMA + IMAD: 1 per 2 DFMA");
        RUN2(0, 1, 1, "2-read DFMA + IMAD: 1 per DFMA");
        RUN2(1, 1, 1, "2-read DFMA + 64-bit register copy: 1 per DFMA");
        RUN2(4, 2, 1, "2-read DFMA + 64-bit zeroing (CS2R): 1 per 2 DFMA");
        RUN2(5, 2, 1, "2-read DFMA + 64-bit select (FSEL pair): 1 per 2 DFMA");
        RUN2(6, 2, 1, "2-read DFMA + 64-bit immediate move: 1 per 2 DFMA");
    }
    return 0;
}
