// One-shot all-reduce of the sufficient statistics over NVLink peer memory (NVSwitch: every peer at full
// bandwidth), replacing ncclAllReduce for this latency-bound exchange (n_c*106 + 2 doubles = 6.8 KB at n_c = 8).
//
// Every rank owns a mailbox  data[2][world][cap] + flags[2][world]  in its own HBM; all ranks map all mailboxes
// through CUDA IPC.  One exchange with sequence number seq (parity p = seq & 1):
//   push   : rank r stores its vector into data[p][r][:] of EVERY rank (plain st.global on the mapped peer
//            pointers), __threadfence_system(), and the last block to finish publishes flags[p][r] = seq on
//            every rank;
//   gather : each rank waits until flags[p][0..world) == seq in its OWN mailbox and adds the world vectors in
//            rank order -> the same bits on every rank (deterministic, no broadcast needed).
// Two parities suffice: a rank can only reach exchange k+2 after it has seen every peer's push k+1, which a
// peer issues after it finished reading exchange k.  One process per GPU, so the waiting kernels of different
// ranks always run on different devices.
#include "common.cuh"

namespace mtm {

__global__ void p2p_push_kernel(const double *__restrict__ src, int count, P2PView v, unsigned int seq,
                                unsigned int *done_counter) {
    const int parity = seq & 1u;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) {
        const double val = src[i];
        const size_t off = ((size_t)parity * v.world + v.rank) * v.cap + i;
        for (int q = 0; q < v.world; ++q) v.data[q][off] = val;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int ticket = atomicAdd(done_counter, 1u);
        if (ticket == gridDim.x - 1) {  // every block's stores are fenced: publish
            __threadfence_system();
            for (int q = 0; q < v.world; ++q)
                *reinterpret_cast<volatile unsigned int *>(&v.flags[q][parity * v.world + v.rank]) = seq;
            *done_counter = 0u;
        }
    }
}

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void p2p_gather_kernel(double *__restrict__ dst, int count, P2PView v, unsigned int seq, Ctrl *ctrl, int R) {
    const int parity = seq & 1u;
    __shared__ int s_ok;
    if (threadIdx.x == 0) s_ok = 1;
    __syncthreads();
    if (threadIdx.x < v.world) {
        volatile unsigned int *f = reinterpret_cast<volatile unsigned int *>(&v.flags[v.rank][parity * v.world + threadIdx.x]);
        const unsigned long long t0 = global_ns();
        unsigned int spins = 0;
        while (*f != seq) {
            // a peer never arrived within the wall-clock budget: give up loudly instead of hanging the device
            if ((++spins & 1023u) == 0 && global_ns() - t0 > v.timeout_ns) {
                s_ok = 0;
                break;
            }
        }
    }
    __syncthreads();
    __threadfence_system();
    if (!s_ok) {
        // every fit of a batch learns about it: finalize_kernel stops the loop on this bit (MTM_EXIT_ERROR)
        if (blockIdx.x == 0)
            for (int r = threadIdx.x; r < (R > 1 ? R : 1); r += blockDim.x) atomicOr(&ctrl[r].status, MTM_STATUS_COMM_TIMEOUT);
        return;
    }
    const double *mine = v.data[v.rank] + (size_t)parity * v.world * v.cap;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < v.world; ++r) s += __ldcg(mine + (size_t)r * v.cap + i);
        dst[i] = s;
    }
}

cudaError_t launch_p2p_allreduce(double *buf, int count, const P2PView &v, unsigned int seq, unsigned int *done_counter,
                                 Ctrl *ctrl, int R, cudaStream_t s) {
    const int threads = 128;
    const int blocks = (count + threads - 1) / threads;
    p2p_push_kernel<<<blocks, threads, 0, s>>>(buf, count, v, seq, done_counter);
    // every block of the gather waits on the flags itself, so any grid size is safe
    p2p_gather_kernel<<<blocks < 32 ? blocks : 32, threads, 0, s>>>(buf, count, v, seq, ctrl, R);
    return cudaGetLastError();
}

}  // namespace mtm
