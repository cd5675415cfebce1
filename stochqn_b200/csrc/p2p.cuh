// p2p.cuh - small-payload all-reduce over NVLink peer memory, fused into the kernel that produces the values.
//
// A sharded optimizer step exchanges a few hundred bytes per reduction phase (4m+2 partial dots, 2 curvature
// dots, k Fisher dots).  A library collective costs three launches (reduce partials, ncclAllReduce, solve) and
// its own latency; here the ONE CTA that has just summed the partials writes them straight into every peer's
// mailbox (st.global on cudaIpc-mapped peer pointers, NVSwitch gives every peer full bandwidth), raises a flag
// per peer, waits for the peers' flags in its own mailbox and adds the records in rank order - so every rank
// ends with bit-identical sums and carries on with the m x m solve in the same launch.
//
// Mailbox of one rank (device memory, zero-initialised, IPC-shared):
//     flags[2][kMaxWorld]           u64: sequence number of the last record rank r has completed in parity p
//     data [2][kMaxWorld][kBoxCap]  f64: rank r's record of that exchange
// Double-buffered by the parity of the exchange sequence number: a rank can only start exchange k+2 (same
// parity as k) after every peer has posted its flag for k+1, i.e. after every peer has finished reading k.
// All exchanges of one communicator must be issued in the same order on every rank and stream-ordered
// on each rank (they are: one stream per workspace, same call sequence on every rank).
// One rank per GPU only: the waiting kernels of all ranks must be co-resident.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sqn {

constexpr int kMaxWorld = 16;
constexpr int kBoxCap = 2048;                     // doubles per rank per exchange
constexpr size_t kBoxFlagBytes = 2 * kMaxWorld * sizeof(unsigned long long);
constexpr size_t kBoxBytes = kBoxFlagBytes + (size_t) 2 * kMaxWorld * kBoxCap * sizeof(double);

struct PeerArgs {
    int rank = 0, world = 0;                      // world <= 1: no exchange
    unsigned long long seq = 0;                   // sequence number of THIS exchange (same on every rank, starts at 1)
    unsigned char* box[kMaxWorld] = {};           // box[r]: rank r's mailbox as mapped into this process
};

__device__ __forceinline__ unsigned long long* box_flag(unsigned char* box, int parity, int r)
{
    return reinterpret_cast<unsigned long long*>(box) + parity * kMaxWorld + r;
}
__device__ __forceinline__ double* box_data(unsigned char* box, int parity, int r)
{
    return reinterpret_cast<double*>(box + kBoxFlagBytes) + ((size_t) parity * kMaxWorld + r) * kBoxCap;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(double* p, double v)
{
    asm volatile("st.relaxed.sys.global.f64 [%0], %1;" :: "l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys(const double* p)
{
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// In-place sum over ranks of vals[0..count) (global or shared memory, visible to the whole CTA).  Called by ALL
// threads of ONE CTA.  Returns false when a peer did not show up within ~20 s (the caller reports it through its
// status word instead of hanging the GPU).
__device__ __forceinline__ bool p2p_allreduce_cta(const PeerArgs& pa, double* vals, int count)
{
    __shared__ int timed_out;
    const int par = (int) (pa.seq & 1ull);
    const int nthreads = blockDim.x;
    if (threadIdx.x == 0) timed_out = 0;
    __syncthreads();                                       // vals written by other threads of the CTA are visible
    for (int t = threadIdx.x; t < pa.world * count; t += nthreads) {
        const int r = t / count, p = t - r * count;
        st_relaxed_sys(box_data(pa.box[r], par, pa.rank) + p, vals[p]);
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < pa.world) {
        st_release_sys(box_flag(pa.box[threadIdx.x], par, pa.rank), pa.seq);
        const unsigned long long* f = box_flag(pa.box[pa.rank], par, threadIdx.x);
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys(f) < pa.seq) {
            if (global_timer_ns() - t0 > 20000000000ull) { timed_out = 1; break; }
        }
    }
    __syncthreads();
    for (int p = threadIdx.x; p < count; p += nthreads) {
        double v = 0;
        for (int r = 0; r < pa.world; ++r) v += ld_relaxed_sys(box_data(pa.box[pa.rank], par, r) + p);
        vals[p] = v;
    }
    __syncthreads();
    return timed_out == 0;
}

}  // namespace sqn
