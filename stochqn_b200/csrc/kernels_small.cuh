// kernels_small.cuh - the whole take_step of oLBFGS / SQN in ONE launch, for latency-bound problem sizes.
//
// For small n (the reference's own CPU-sized configurations: n = 1001, 4097 ...) the three launches K1 -> K2 -> K3
// are dominated by launch latency and by the host waiting for K2's flag, not by memory traffic: every n-vector is
// a few KB and lives in L2.  ks_step does the same arithmetic in one cooperative launch:
//
//   phase 1  every CTA owns a contiguous slice of the elements; its warps deal the (up to 4m+2) inner products of
//            K1's record among themselves, one product per warp at a time over the slice (g and y_c are re-read from
//            L1), fp64 accumulation, one partial record per CTA; the oLBFGS copy grad_prev <- g rides along
//   barrier  one grid-wide barrier (all CTAs are resident: cooperative launch, grid <= SM count)
//   phase 2  EVERY CTA sums the partial records in the same fixed order and solves the m x m compact-form system
//            redundantly in its own shared memory (solve_cta, kernels.cuh) - no second barrier, no broadcast; CTA 0
//            alone folds the pending Gram column into the global Gram state, writes the coefficient block the
//            exact-norm fallback reads, and publishes the status word to the host
//   phase 3  each thread combines its elements exactly as K3 does (same FMA order, so the two paths agree to the
//            last bit given the same coefficients): d = (gamma*g + sum_j a_j s_j) + sum_j (gamma b_j) y_j,
//            x -= step*d, and the optimizer-specific epilogue
//
// Reference lines replaced: stochqn.c:663-708 (two-loop), 802-840 (take_step), 996, 1006-1007, 1067.
#pragma once

namespace sqn {

struct SmallArgs {
    int msize, used, pend, new_slot;
    long long n;
    size_t ld;
    unsigned long long bar_target;       // value the barrier counter reaches when every CTA of this launch has arrived
};

// grid-wide barrier on a monotonically increasing counter (cooperative launch guarantees co-residency)
__device__ __forceinline__ void grid_barrier(unsigned long long* counter, unsigned long long target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1ull);
        unsigned long long v;
        do {
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(counter) : "memory");
        } while (v < target);
        __threadfence();
    }
    __syncthreads();
}

template <typename T>
__device__ __forceinline__ double slice_dot(const T* __restrict__ a, const T* __restrict__ b, long long e0, long long e1, int lane)
{
    // 16 independent loads in flight per lane in EVERY trip, the ragged end included (predicated, not a scalar tail loop:
    // at these sizes the kernel is a chain of load latencies, and a one-element-per-trip tail was most of it)
    double acc = 0;
    for (long long base = e0; base < e1; base += 8 * 32) {
        T av[8], bv[8];
        #pragma unroll
        for (int u = 0; u < 8; ++u) {
            const long long i = base + lane + u * 32;
            const bool ok = i < e1;
            av[u] = ok ? a[i] : (T) 0;
            bv[u] = ok ? b[i] : (T) 0;
        }
        #pragma unroll
        for (int u = 0; u < 8; ++u) acc = fma((double) av[u], (double) bv[u], acc);
    }
    return warp_sum(acc);
}

// d_i = (gamma*g_i + sum_j a_j s_ji) + sum_j (gamma b_j) y_ji with K3's FMA order; the row loads are issued eight pairs at
// a time ahead of the FMA chains (a plain loop over a run-time row count serialises 2*used load latencies)
template <typename T>
__device__ __forceinline__ T combine_direction_at(T gi, const T* S, const T* __restrict__ Y, size_t ld, long long i,
                                                  int used, int m, T gamma, const double* coef_s)
{
    T p0 = gamma * gi, p1 = (T) 0;
    for (int r0 = 0; r0 < used; r0 += 8) {
        T sv[8], yv[8];
        #pragma unroll
        for (int q = 0; q < 8; ++q) {
            const bool ok = r0 + q < used;
            sv[q] = ok ? S[(size_t) (r0 + q) * ld + i] : (T) 0;
            yv[q] = ok ? Y[(size_t) (r0 + q) * ld + i] : (T) 0;
        }
        #pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (r0 + q < used) {
                p0 = fma((T) coef_s[r0 + q], sv[q], p0);
                p1 = fma((T) coef_s[m + r0 + q], yv[q], p1);
            }
        }
    }
    return p0 + p1;
}

template <typename T>
__device__ __forceinline__ T combine_direction(const T* __restrict__ g, const T* S, const T* __restrict__ Y, size_t ld, long long i,
                                               int used, int m, T gamma, const double* coef_s)
{
    return combine_direction_at<T>(g[i], S, Y, ld, i, used, m, gamma, coef_s);
}

// Launched either as ONE CTA of 1024 threads with an ordinary launch (n <= kOneCtaN = 2048: no grid barrier at all) or as a
// cooperative grid of 256-thread CTAs (one barrier).
constexpr int kOneCtaThreads = 1024;
constexpr long long kOneCtaN = 2048;

template <typename T, int MODE>
__global__ void __launch_bounds__(kOneCtaThreads)
ks_step(SmallArgs K, SolveArgs A, const T* g, T* gout, T* S, const T* __restrict__ Y,
        T* __restrict__ x, T* __restrict__ x_sum, T* __restrict__ grad_prev, T step,
        double* __restrict__ partials, double* __restrict__ SY, double* __restrict__ YY, double* __restrict__ SS,
        double* __restrict__ coef, int* __restrict__ status_dev, volatile int* status_host, volatile double* info_host,
        volatile unsigned long long* seq_host, unsigned long long* bar)
{
    const int m = K.msize, used = K.used, c = K.pend;
    const int P = 4 * m + 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nthr = (int) blockDim.x, nwarps = nthr >> 5;
    const long long per = (K.n + gridDim.x - 1) / gridDim.x;
    const long long e0 = (long long) blockIdx.x * per;
    const long long e1 = e0 + per < K.n ? e0 + per : K.n;
    __shared__ double sums_s[4 * kMaxMem + 2];
    __shared__ double coef_s[2 * kMaxMem + 3];
    __shared__ SolveShared sh;

    // ---- phase 1: this CTA's partial record --------------------------------------------------------------
    double* rec = partials + (size_t) blockIdx.x * P;
    for (int p = threadIdx.x; p < P; p += nthr) rec[p] = 0.0;
    __syncthreads();
    const int nd = 2 * used + 1 + (c >= 0 ? 2 * used + 1 : 0);
    const T* yc = c >= 0 ? Y + (size_t) c * K.ld : nullptr;
    const T* sc = c >= 0 ? S + (size_t) c * K.ld : nullptr;
    for (int d = warp; d < nd; d += nwarps) {
        const T *a, *b;
        int idx;
        if (d < used)              { a = S + (size_t) d * K.ld;                  b = g;  idx = d; }
        else if (d < 2 * used)     { a = Y + (size_t) (d - used) * K.ld;         b = g;  idx = m + (d - used); }
        else if (d == 2 * used)    { a = g;                                      b = g;  idx = 4 * m; }
        else if (d < 3 * used + 1) { a = S + (size_t) (d - 2 * used - 1) * K.ld; b = yc; idx = 2 * m + (d - 2 * used - 1); }
        else if (d < 4 * used + 1) { a = Y + (size_t) (d - 3 * used - 1) * K.ld; b = yc; idx = 3 * m + (d - 3 * used - 1); }
        else                       { a = sc;                                     b = sc; idx = 4 * m + 1; }
        const double v = slice_dot(a, b, e0, e1, lane);
        if (lane == 0) rec[idx] = v;
    }
    if (grad_prev) for (long long i = e0 + threadIdx.x; i < e1; i += nthr) grad_prev[i] = g[i];

    if (gridDim.x > 1) grid_barrier(bar, K.bar_target);
    else __syncthreads();

    // ---- phase 2: every CTA reduces the records in the same order and solves ------------------------------
    {
        const int nb = (int) gridDim.x;
        for (int p = warp; p < P; p += nwarps) {
            double v = 0;
            for (int b = lane; b < nb; b += 32) v += __ldcg(partials + (size_t) b * P + p);
            v = warp_sum(v);
            if (lane == 0) sums_s[p] = v;
        }
        __syncthreads();
    }
    const int st = solve_cta(A, sums_s, SY, YY, SS, sh, coef_s, blockIdx.x == 0, true, nthr);
    if (blockIdx.x == 0) {
        for (int j = threadIdx.x; j < 2 * m + 3; j += nthr) coef[j] = coef_s[j];
        __syncthreads();
        if (threadIdx.x == 0) publish_status(st, coef_s, m, status_dev, status_host, info_host, seq_host, A.seq);
    }
    if (st != ST_ACCEPT) return;

    // ---- phase 3: combine + update (arithmetic of k3_combine) ------------------------------------------
    const T gamma = (T) coef_s[2 * m];
    const T nstep = -step;
    for (long long i = e0 + threadIdx.x; i < e1; i += nthr) {
        T d = combine_direction<T>(g, S, Y, K.ld, i, used, m, gamma, coef_s);
        const T xv = fma(nstep, d, x[i]);
        x[i] = xv;
        if constexpr (MODE == MODE_OLBFGS) {
            d = nstep * d;
            S[(size_t) K.new_slot * K.ld + i] = d;
        } else {
            x_sum[i] = x_sum[i] + xv;
        }
        if (gout) gout[i] = d;
    }
}

}  // namespace sqn
