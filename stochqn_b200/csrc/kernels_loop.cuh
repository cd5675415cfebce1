// kernels_loop.cuh - optimizer steps whose control flow lives ON THE DEVICE (no host round trip).
//
// The free-mode ABI (run_oLBFGS / run_SQN / run_adaQN) must hand the next task back to its caller, so every call
// waits for a flag the device publishes (accept / reject of the direction, the two curvature dots).  When the caller
// is the library itself - the guided request loop over the bundled callbacks, stochqn_b200_fit_batches - nobody needs
// those answers on the host: the ring-buffer counters (mem_used, mem_st_ix, the pending Gram column) move into a small
// device record, every kernel reads them from there, takes the reference's decisions itself and leaves the outcome
// for the next kernel.  The host only enqueues a fixed kernel sequence per mini-batch and reads the counters back
// when it is asked for them (once per call of stochqn_b200_fit_batches, not per request).
//
//   kl_step   oLBFGS / SQN take_step (stochqn.c:802-840 + 663-708) in ONE launch: dots, solve, accept test (with the
//             exact-norm route inside), update, and the bookkeeping of a rejected direction (flush, quirk Q7)
//   kl_pair   oLBFGS pair call (stochqn.c:915-926, 883-900): y = g - g_prev (+ y_reg s), curvature test, quirk Q1
//             on rejection, ring advance on acceptance; a no-op when the step before it was rejected (the reference
//             then asks for a gradient on a NEW batch instead, stochqn.c:1010-1020)
//   kl_ada    adaQN take_step (stochqn.c:802-840 with 720-783) in ONE launch: accumulator update, Fisher ring write,
//             the two reduction phases of the compact form with a diagonal H0 (kernels_adaqn.cuh), update (512-thread CTAs)
//
// kl_step / kl_pair are cooperative grids of 256-thread CTAs (or one 1024-thread CTA for n <= 2048) in which every CTA owns a
// contiguous slice of the elements; partial records are summed by every CTA in the same fixed order, and every CTA
// solves the m x m system redundantly in its own shared memory, so no broadcast is needed after a grid barrier.
// Arithmetic of the update is that of K3 / KA3 (same FMA order), so the iterates agree with the host-driven routes to
// the last bit given the same coefficients.
#pragma once

#include "mn_small.cuh"

namespace sqn {

struct LoopState {
    int used, st_ix, pend, skip_pair;           // ring state; skip_pair: the last oLBFGS step was rejected
    int fisher_used, fisher_st, last_status, last_info;      // last_info: info_enum code (200..203) of the last call
    unsigned long long n_info[4];               // calls by info code: 200 ok, 201 func_increased, 202 curvature, 203 nan direction
    unsigned long long calls;                   // optimizer calls the reference's loop would have made
    unsigned long long x_changed;               // steps that updated x
};

struct LoopArgs {
    int msize, check_nan, mode, pad;
    long long n;
    size_t ld;
    double h0, limit;                           // hess_init (oLBFGS) or 0; 1e3 * n
    double min_curvature, y_reg;
};

// Grid barrier on a monotonically increasing counter that is a multiple of gridDim.x whenever no barrier is in
// progress (every launch that uses it has the same grid and runs a whole number of barriers): the target is derived
// from the ticket, so the host does not have to know how many barriers a launch will execute (the exact-norm route
// adds one).  Cooperative launch guarantees co-residency.
__device__ __forceinline__ void grid_barrier_auto(unsigned long long* counter)
{
    __syncthreads();
    if (gridDim.x > 1 && threadIdx.x == 0) {
        __threadfence();
        const unsigned long long old = atomicAdd(counter, 1ull);
        const unsigned long long target = (old / gridDim.x + 1ull) * gridDim.x;
        unsigned long long v;
        do {
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(counter) : "memory");
        } while (v < target);
        __threadfence();
    }
    __syncthreads();
}

// CTA sum of two per-thread accumulators into out[0..1] (global), fixed order, any block size that is a multiple of 32
__device__ __forceinline__ void block_sum2_to(double a, double b, double* __restrict__ out)
{
    __shared__ double red2[32][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (int) (blockDim.x >> 5);
    a = warp_sum(a);
    b = warp_sum(b);
    if (lane == 0) { red2[warp][0] = a; red2[warp][1] = b; }
    __syncthreads();
    if (threadIdx.x < 2) {
        double v = 0;
        for (int w = 0; w < nw; ++w) v += red2[w][threadIdx.x];
        out[threadIdx.x] = v;
    }
}

// sum the 2-value records of all CTAs in CTA order (every CTA does it: same result everywhere)
__device__ __forceinline__ void reduce2_all(const double* __restrict__ rec2, double* out_s /* shared, 2 */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp < 2) {
        double v = 0;
        for (unsigned b = lane; b < gridDim.x; b += 32) v += __ldcg(rec2 + (size_t) b * 2 + warp);
        v = warp_sum(v);
        if (lane == 0) out_s[warp] = v;
    }
    __syncthreads();
}

template <typename T, int MODE>
__global__ void __launch_bounds__(kOneCtaThreads)
kl_step(LoopArgs A, LoopState* __restrict__ st, const T* g, T* gout, T* S, const T* __restrict__ Y,
        T* __restrict__ x, T* __restrict__ x_sum, T* __restrict__ grad_prev, T step,
        double* __restrict__ partials, double* __restrict__ rec2, double* __restrict__ SY, double* __restrict__ YY,
        double* __restrict__ SS, double* __restrict__ coef, unsigned long long* bar)
{
    const int m = A.msize;
    const int used = st->used, slot = st->st_ix, c = st->pend;         // read before anybody may rewrite them (CTA 0, at the very end)
    const int P = 4 * m + 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nthr = (int) blockDim.x, nwarps = nthr >> 5;
    const long long per = (A.n + gridDim.x - 1) / gridDim.x;
    const long long e0 = (long long) blockIdx.x * per;
    const long long e1 = e0 + per < A.n ? e0 + per : A.n;
    __shared__ double sums_s[4 * kMaxMem + 2];
    __shared__ double coef_s[2 * kMaxMem + 3];
    __shared__ double two_s[2];
    __shared__ SolveShared sh;

    // ---- phase 1: this CTA's partial record (ks_step, kernels_small.cuh) ----------------------------------
    double* rec = partials + (size_t) blockIdx.x * P;
    for (int p = threadIdx.x; p < P; p += nthr) rec[p] = 0.0;
    __syncthreads();
    const int nd = 2 * used + 1 + (c >= 0 ? 2 * used + 1 : 0);
    const T* yc = c >= 0 ? Y + (size_t) c * A.ld : nullptr;
    const T* sc = c >= 0 ? S + (size_t) c * A.ld : nullptr;
    for (int d = warp; d < nd; d += nwarps) {
        const T *a, *b;
        int idx;
        if (d < used)              { a = S + (size_t) d * A.ld;                  b = g;  idx = d; }
        else if (d < 2 * used)     { a = Y + (size_t) (d - used) * A.ld;         b = g;  idx = m + (d - used); }
        else if (d == 2 * used)    { a = g;                                      b = g;  idx = 4 * m; }
        else if (d < 3 * used + 1) { a = S + (size_t) (d - 2 * used - 1) * A.ld; b = yc; idx = 2 * m + (d - 2 * used - 1); }
        else if (d < 4 * used + 1) { a = Y + (size_t) (d - 3 * used - 1) * A.ld; b = yc; idx = 3 * m + (d - 3 * used - 1); }
        else                       { a = sc;                                     b = sc; idx = 4 * m + 1; }
        const double v = slice_dot(a, b, e0, e1, lane);
        if (lane == 0) rec[idx] = v;
    }
    if (grad_prev) for (long long i = e0 + threadIdx.x; i < e1; i += nthr) grad_prev[i] = g[i];
    grid_barrier_auto(bar);

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
    SolveArgs SA;
    SA.msize = m; SA.used = used; SA.oldest = (slot == used) ? 0 : slot; SA.pend = c; SA.nblocks = 0; SA.do_solve = 1;
    SA.check_nan = A.check_nan; SA.h0 = A.h0; SA.limit = A.limit; SA.seq = 0;
    int status = solve_cta(SA, sums_s, SY, YY, SS, sh, coef_s, blockIdx.x == 0, true, nthr);
    const T gamma = (T) coef_s[2 * m];
    const T nstep = -step;

    auto direction = [&](long long i) -> T { return combine_direction<T>(g, S, Y, A.ld, i, used, m, gamma, coef_s); };

    bool d_in_g = false;
    if (status == ST_NEED_EXACT_NORM) {
        // the bound could not certify ||d|| <= 1e3*n: materialise d in `g`, measure it exactly, then decide - the
        // reference's check-before-update order (stochqn.c:825-838).  Every CTA takes this branch together.
        double a_dd = 0, a_bad = 0;
        T* gw = const_cast<T*>(g);
        for (long long i = e0 + threadIdx.x; i < e1; i += nthr) {
            const T d = direction(i);
            const double de = (double) d;
            a_dd = fma(de, de, a_dd);
            if (!isfinite(de)) a_bad += 1.0;
            gw[i] = d;
        }
        block_sum2_to(a_dd, a_bad, rec2 + (size_t) blockIdx.x * 2);
        grid_barrier_auto(bar);
        reduce2_all(rec2, two_s);
        status = (two_s[1] > 0 || !(sqrt(two_s[0]) <= A.limit)) ? ST_REJECT_NONFINITE : ST_ACCEPT;
        d_in_g = true;
    }

    if (status == ST_ACCEPT) {
        // ---- phase 3: combine + update ---------------------------------------------------------------------
        for (long long i = e0 + threadIdx.x; i < e1; i += nthr) {
            T d = d_in_g ? g[i] : direction(i);
            const T xv = fma(nstep, d, x[i]);
            x[i] = xv;
            if constexpr (MODE == MODE_OLBFGS) {
                d = nstep * d;
                S[(size_t) slot * A.ld + i] = d;
            } else {
                x_sum[i] = x_sum[i] + xv;
            }
            if (gout) gout[i] = d;
        }
    } else if constexpr (MODE == MODE_AVG) {
        for (long long i = e0 + threadIdx.x; i < e1; i += nthr) x_sum[i] = x_sum[i] + x[i];      // quirk Q7 (stochqn.c:1067)
    }

    if (blockIdx.x == 0) {
        for (int j = threadIdx.x; j < 2 * m + 3; j += nthr) coef[j] = coef_s[j];
        if (threadIdx.x == 0) {
            st->last_status = status;
            st->calls += 1;
            if (status == ST_ACCEPT) {
                st->pend = -1;
                st->skip_pair = 0;
                st->n_info[0] += 1;
                st->x_changed += 1;
                st->last_info = 200;
            } else {                                    // flush_bfgs_mem (stochqn.c:554-558), search_direction_was_nan
                st->used = 0; st->st_ix = 0; st->pend = -1;
                st->skip_pair = 1;
                st->n_info[3] += 1;
                st->last_info = 203;
            }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(kOneCtaThreads)
kl_pair(LoopArgs A, LoopState* __restrict__ st, const T* __restrict__ g, const T* __restrict__ g_prev, T* S, T* Y,
        double* __restrict__ rec2, double* __restrict__ SY, double* __restrict__ YY, double* __restrict__ SS,
        unsigned long long* bar)
{
    if (st->skip_pair) return;                          // the step was rejected: the reference asks for a new batch instead
    const int m = A.msize;
    const int used = st->used, slot = st->st_ix;
    const int nthr = (int) blockDim.x;
    const long long per = (A.n + gridDim.x - 1) / gridDim.x;
    const long long e0 = (long long) blockIdx.x * per;
    const long long e1 = e0 + per < A.n ? e0 + per : A.n;
    T* s = S + (size_t) slot * A.ld;
    T* y = Y + (size_t) slot * A.ld;
    const T y_reg = (T) A.y_reg;
    __shared__ double two_s[2];
    double a_sy = 0, a_ss = 0;
    for (long long i = e0 + threadIdx.x; i < e1; i += nthr) {
        const T sv = s[i];
        T t = g[i] - g_prev[i];
        if (y_reg > (T) 0) t = fma(y_reg, sv, t);
        y[i] = t;
        a_sy = fma((double) sv, (double) t, a_sy);
        a_ss = fma((double) sv, (double) sv, a_ss);
    }
    block_sum2_to(a_sy, a_ss, rec2 + (size_t) blockIdx.x * 2);
    grid_barrier_auto(bar);
    reduce2_all(rec2, two_s);
    // check_min_curvature (stochqn.c:883-900): the division in T, as the reference's (and as the host route)
    const bool reject = A.min_curvature > 0 && ((T) two_s[0] / (T) two_s[1]) <= (T) A.min_curvature;
    if (reject) {
        for (long long i = e0 + threadIdx.x; i < e1; i += nthr) { s[i] = (T) 0; y[i] = (T) 0; }    // quirk Q1
        if (blockIdx.x == 0) {
            for (int j = threadIdx.x; j < m; j += nthr) {
                SY[j * m + slot] = 0; SY[slot * m + j] = 0;
                YY[j * m + slot] = 0; YY[slot * m + j] = 0;
            }
            if (threadIdx.x == 0) SS[slot] = 0;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->calls += 1;
        st->last_info = reject ? 202 : 200;
        if (reject) st->n_info[2] += 1;
        else {
            st->n_info[0] += 1;
            st->st_ix = (slot + 1) % m;                 // incr_bfgs_counters (stochqn.c:569-573)
            st->used = used + 1 >= m ? m : used + 1;
            st->pend = slot;
        }
    }
}

// Grid barrier (cooperative launch: all CTAs resident).  bar[0] counts arrivals and only ever grows; it is a multiple of
// gridDim.x whenever no barrier is in progress (every launch that uses it has the same grid), so the generation a CTA
// waits for follows from its ticket and the host need not know how many barriers a launch executes (rejections and the
// exact-norm route change it).  The CTA that arrives last publishes the generation in bar[32] - another 128-byte line,
// so the waiters' polling does not queue up behind the arrivals at the same L2 atomic unit.
constexpr int kFitBarWords = 64;
__device__ __forceinline__ void fit_barrier(unsigned long long* bar)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned long long old = atomicAdd(bar, 1ull);
        const unsigned long long gen = old / gridDim.x + 1ull;
        unsigned long long* flag = bar + 32;
        if ((old + 1ull) % gridDim.x == 0) {
            asm volatile("st.release.gpu.global.u64 [%0], %1;" :: "l"(flag), "l"(gen) : "memory");
        } else {
            unsigned long long v;
            do {
                asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
            } while (v < gen);
        }
    }
    __syncthreads();
}

// The same barrier for PERSISTENT kernels whose counter the host zeroes before the launch: the k-th barrier of a CTA is passed
// when the counter reaches k * gridDim.x, so the arrival is a fire-and-forget `red` (no ticket to wait for: one L2 round trip
// less on the critical path of every barrier) and the waiters poll the counter itself.
struct RunBarrier {
    unsigned long long* bar;
    unsigned long long passed;
    __device__ __forceinline__ void operator()()
    {
        __syncthreads();
        passed += gridDim.x;
        if (threadIdx.x == 0) {
            asm volatile("red.release.gpu.global.add.u64 [%0], 1;" :: "l"(bar) : "memory");
            unsigned long long v;
            do {
                asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(bar) : "memory");
            } while (v < passed);
        }
        __syncthreads();
    }
};

// =========================================================================================
// kl_ada: the adaQN take_step (stochqn.c:802-840 with 720-783; kernels_adaqn.cuh for the compact form with the
// diagonal H0 = diag(h)) in ONE cooperative launch, decisions on the device.  KA1 -> KAu -> KA2 -> KAa -> KA3 become
// three sweeps of a CTA's own contiguous slice of the elements separated by two grid barriers; every CTA sums the
// partial records in the same order and performs both small solves redundantly in its own shared memory.
//   A  G <- accumulate(g), Fisher ring row <- g, p = S'g, pending Gram column, sum h^2          -- barrier
//   B  u = R^-1 p ;  w = Y'[h.(Yu - g)], sum (h.(Yu-g))^2                                       -- barrier
//   C  a = R^-T (D u + w), bound on ||d||, accept / reject (exact-norm route: + one barrier);
//      d = h.(g + Y b) + S a ;  x -= step*d ;  x_sum += x       (rejected: x_sum += x, flush; quirk Q7)
// Element arithmetic is that of ka1_dots / ka2_wdots / ka3_combine (same FMA order).
// =========================================================================================
constexpr int kAdaThreads = 512;

// Sum N (a power of two <= 32) per-lane values over the warp with N - 1 + log2(32 / N) shuffles of doubles instead of 5 N: at
// every step a lane keeps the half of the values its lane bit selects and hands the other half to its partner.  Afterwards
// lane l holds, in the return value, the warp total of value number l / (32 / N).  Fixed order: deterministic.
template <int N>
__device__ __forceinline__ double warp_reduce_multi(double (&v)[N])
{
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    int o = 16;
    #pragma unroll
    for (int half = N / 2; half >= 1; half >>= 1, o >>= 1) {
        const bool upper = (lane & o) != 0;
        #pragma unroll
        for (int j = 0; j < half; ++j) {
            const double send = upper ? v[j] : v[j + half];
            const double keep = upper ? v[j + half] : v[j];
            v[j] = keep + __shfl_xor_sync(full, send, o);
        }
    }
    double r = v[0];
    #pragma unroll
    for (; o >= 1; o >>= 1) r += __shfl_xor_sync(full, r, o);
    return r;
}

struct AdaLoopArgs {
    int msize, check_nan, fisher_size, pad;
    long long n;
    size_t ld;
    double limit;                                // 1e3 * n
    double scal_reg, rmsprop_weight;
};

struct AdaShared {
    double Rm[kMaxMem][kMaxMem + 1];
    double pv[kMaxMem], u[kMaxMem], w[kMaxMem], ssv[kMaxMem];
    int status;
};

// sum entry-major records: entry p of CTA b is rec[p * G + b]; one warp per entry (three at a time), lanes over CTAs
__device__ __forceinline__ void reduce_entry_major(const double* __restrict__ rec, int P, int G, double* __restrict__ sums, int nwarps)
{
    // A warp sums three entries at a time, lanes over CTAs, and ALL their loads (3 x up to 5 x 32 records per trip) are issued before
    // the first add: for the usual 4m + 2 = 42 entries on 16 warps the whole reduction is one L2 round trip long.
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int p0 = warp; p0 < P; p0 += 3 * nwarps) {
        double acc[3] = {0.0, 0.0, 0.0};
        for (int r0 = 0; r0 < G; r0 += 5 * 32) {
            double v[3][5];
            #pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int p = p0 + q * nwarps;
                #pragma unroll
                for (int u = 0; u < 5; ++u) {
                    const int r = r0 + u * 32 + lane;
                    v[q][u] = (p < P && r < G) ? __ldcg(rec + (size_t) p * G + r) : 0.0;
                }
            }
            #pragma unroll
            for (int q = 0; q < 3; ++q)
                #pragma unroll
                for (int u = 0; u < 5; ++u) acc[q] += v[q][u];
        }
        #pragma unroll
        for (int q = 0; q < 3; ++q) {
            const int p = p0 + q * nwarps;
            if (p < P) {                                     // warp-uniform
                const double t = warp_sum(acc[q]);
                if (lane == 0) sums[p] = t;
            }
        }
    }
    __syncthreads();
}

// the ring / Fisher counters and the tallies of a device-side adaQN loop, carried in registers by every CTA
struct AdaRing {
    int used, slot, pend, f_used, f_st;
    int last_status, last_info;
    unsigned long long n_ok, n_nan, calls, x_changed;
};

__device__ __forceinline__ AdaRing ada_ring_load(const LoopState* st)
{
    AdaRing R;
    R.used = st->used; R.slot = st->st_ix; R.pend = st->pend; R.f_used = st->fisher_used; R.f_st = st->fisher_st;
    R.last_status = st->last_status; R.last_info = st->last_info;
    R.n_ok = 0; R.n_nan = 0; R.calls = 0; R.x_changed = 0;
    return R;
}

__device__ __forceinline__ void ada_ring_store(LoopState* st, const AdaRing& R)       // one thread of one CTA, after the last barrier
{
    st->used = R.used; st->st_ix = R.slot; st->pend = R.pend; st->fisher_used = R.f_used; st->fisher_st = R.f_st;
    st->last_status = R.last_status; st->last_info = R.last_info;
    st->n_info[0] += R.n_ok; st->n_info[3] += R.n_nan; st->calls += R.calls; st->x_changed += R.x_changed;
}

// Whole-grid function: every CTA of a cooperative grid of kAdaThreads-thread CTAs calls it with the same arguments and the same
// ring state `R` (updated in place, identically everywhere).  `ada_smem`: dynamic shared memory, 2 * ceil(n / grid) elements.
template <typename T, int MMAX, typename Barrier>
__device__ __forceinline__ void kl_ada_body(const AdaLoopArgs& A, AdaRing& R, const T* g, T* gout, T* __restrict__ Gacc,
       const T* __restrict__ S, const T* __restrict__ Y, T* __restrict__ F, T* __restrict__ x, T* __restrict__ x_sum, const T step,
       double* partials, double* rec2, double* SY, double* YY, double* SS, Barrier&& barrier, unsigned long long* trace,
       unsigned char* ada_smem)
{
    const int m = A.msize;
    const int used = R.used, slot = R.slot, c = R.pend;
    auto stamp = [&](int k) {                                // development aid (stochqn_b200_debug_fit_trace)
        if (trace && blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); trace[k] = t; }
    };
    stamp(0);
    const int f_used = R.f_used, f_st = R.f_st;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = (int) gridDim.x, b = (int) blockIdx.x;
    constexpr int NW = kAdaThreads / 32;
    const long long per = (A.n + G - 1) / G;
    const long long e0 = (long long) b * per < A.n ? (long long) b * per : A.n;
    const long long e1 = e0 + per < A.n ? e0 + per : A.n;
    const T scal_reg = (T) A.scal_reg, rmsw = (T) A.rmsprop_weight;
    const bool rms = (rmsw > (T) 0 && rmsw < (T) 1);
    const T w_new = (T) 1 - rmsw;
    const int oldest = (slot == used) ? 0 : slot;
    auto ph = [&](int i) { int s2 = oldest + i; return s2 >= m ? s2 - m : s2; };
    T* frow = A.fisher_size > 0 ? F + (size_t) f_st * A.ld : nullptr;
    const T* yc = c >= 0 ? Y + (size_t) c * A.ld : nullptr;
    const T* sc = c >= 0 ? S + (size_t) c * A.ld : nullptr;

    T* g_s = reinterpret_cast<T*>(ada_smem);                  // this CTA's slice of g and of h = g / sqrt(G + eps): computed once in
    T* h_s = g_s + per;                                        // phase A, read by B and C (the divide + square root are ~100 fp64 issue slots)
    __shared__ double sums1[2 * kMaxMem + 4], sums2[kMaxMem + 1], coef_s[2 * kMaxMem + 3], two_s[2];
    __shared__ double red[NW][2 * MMAX + 4];
    __shared__ AdaShared sh;
    const int P1 = 2 * m + 4, P2 = m + 1;

    // ---- A: accumulator, Fisher row, p, pending column ----
    {
        double a_p[MMAX], a_c[MMAX], a_hh = 0, a_ss = 0, a_yy = 0;
        #pragma unroll
        for (int r = 0; r < MMAX; ++r) { a_p[r] = 0; a_c[r] = 0; }
        for (long long i = e0 + tid; i < e1; i += kAdaThreads) {
            T rv[MMAX];
            const T gv = __ldcg(g + i);                  // (other CTAs of a persistent grid may have written it: L2, not a stale L1 line)
            T Gv = Gacc[i];
            #pragma unroll
            for (int r = 0; r < MMAX; ++r) rv[r] = r < used ? S[(size_t) r * A.ld + i] : (T) 0;
            T ycv = (T) 0, scv = (T) 0;
            if (c >= 0) { ycv = yc[i]; scv = sc[i]; }
            Gv = ada_accumulate<T>(gv, Gv, rmsw, w_new, rms);
            Gacc[i] = Gv;
            if (frow) frow[i] = gv;
            const T h = gv / sqrt(Gv + scal_reg);
            g_s[i - e0] = gv;
            h_s[i - e0] = h;
            a_hh = fma((double) h, (double) h, a_hh);
            if (c >= 0) { a_ss = fma((double) scv, (double) scv, a_ss); a_yy = fma((double) ycv, (double) ycv, a_yy); }
            #pragma unroll
            for (int r = 0; r < MMAX; ++r) {
                if (r < used) {
                    a_p[r] = fma((double) rv[r], (double) gv, a_p[r]);
                    if (c >= 0) a_c[r] = fma((double) rv[r], (double) ycv, a_c[r]);
                }
            }
        }
        {                                                    // warp totals: value q lands in lane q of its group of (up to) 32
            constexpr int NV = 2 * MMAX + 3;
            #pragma unroll
            for (int q0 = 0; q0 < NV; q0 += 32) {
                constexpr int NG = 32;
                double v[NG];
                #pragma unroll
                for (int j = 0; j < NG; ++j) {
                    const int q = q0 + j;
                    v[j] = q < MMAX ? a_p[q < MMAX ? q : 0] : q < 2 * MMAX ? a_c[(q - MMAX) < MMAX && q >= MMAX ? (q - MMAX) : 0]
                         : q == 2 * MMAX ? a_hh : q == 2 * MMAX + 1 ? a_ss : q == 2 * MMAX + 2 ? a_yy : 0.0;
                }
                const double r = warp_reduce_multi<NG>(v);
                if (q0 + lane < NV) red[warp][q0 + lane] = r;
            }
        }
        __syncthreads();
        for (int p = tid; p < P1; p += kAdaThreads) {        // record layout of ka1_dots: [0,m) p, [m,2m) s_j'y_c, 2m: sum h^2, 2m+1: s_c's_c, 2m+2: y_c'y_c
            const int q = p < m ? p : p < 2 * m ? MMAX + (p - m) : 2 * MMAX + (p - 2 * m);
            const bool have = p < m ? p < used : p < 2 * m ? (c >= 0 && p - m < used) : p < 2 * m + 3;
            double v = 0;
            if (have) { for (int w2 = 0; w2 < NW; ++w2) v += red[w2][q]; }
            partials[(size_t) p * G + b] = v;          // (values nobody asked for were summed too: they are zeros or unused)
        }
    }
    stamp(1);
    barrier();
    stamp(2);

    // ---- B: u, then w = Y'[h.(Yu - g)] ----
    reduce_entry_major(partials, P1, G, sums1, NW);
    stamp(3);
    if (b == 0 && c >= 0) {                                  // fold the pending pair's Gram column into the global state
        for (int j = tid; j < used; j += kAdaThreads) SY[j * m + c] = sums1[m + j];
        if (tid == 0) { SS[c] = sums1[2 * m + 1]; YY[c * m + c] = sums1[2 * m + 2]; }
    }
    for (int t = tid; t < used * used; t += kAdaThreads) {
        const int i = t / used, j = t % used;
        const int pi = ph(i), pj = ph(j);
        sh.Rm[i][j] = (pj == c) ? sums1[m + pi] : __ldcg(SY + pi * m + pj);
    }
    for (int i = tid; i < used; i += kAdaThreads) {
        const int pi = ph(i);
        sh.pv[i] = sums1[pi];
        sh.ssv[i] = (pi == c) ? sums1[2 * m + 1] : __ldcg(SS + pi);
    }
    for (int j = tid; j < 2 * m + 3; j += kAdaThreads) coef_s[j] = 0.0;
    __syncthreads();
    if (tid < 32) {                                          // u = R^-1 p, lane i owns row i (ada_stage_and_solve_u)
        const bool live = tid < used;
        const double inv = 1.0 / (live ? sh.Rm[tid][tid] : 1.0);
        double t = live ? sh.pv[tid] : 0.0;
        for (int j = used - 1; j >= 0; --j) {
            const double uj = __shfl_sync(0xffffffffu, t * inv, j);
            if (tid == j) sh.u[tid] = uj;
            if (tid < j) t = fma(-sh.Rm[tid][j], uj, t);
        }
    }
    __syncthreads();
    for (int i = tid; i < used; i += kAdaThreads) coef_s[m + ph(i)] = -sh.u[i];
    __syncthreads();
    stamp(4);

    int status = ST_ACCEPT;
    if (used == 0) {                                         // d = h: its norm is exact (stochqn.c:808-812, 825-835)
        const double hh = sums1[2 * m];
        if (A.check_nan && (!isfinite(hh) || !(sqrt(hh) <= A.limit))) status = ST_REJECT_NONFINITE;
    } else {
        // two elements of the slice per trip: 2 * used row loads in flight per thread (the loop is a chain of L2 round trips - half as
        // many this way); the coefficients u_j are read from shared memory (broadcast) instead of living in registers.  (The same
        // unrolling of phases A and C costs more in spills than it saves: 128 registers per thread at 512 threads.)
        double acc[MMAX], a_tt = 0;
        #pragma unroll
        for (int j = 0; j < MMAX; ++j) acc[j] = 0;
        for (long long i = e0 + tid; i < e1; i += 2 * kAdaThreads) {
            const long long i1 = i + kAdaThreads;
            const bool ok1 = i1 < e1;
            T y0[MMAX], y1[MMAX];
            #pragma unroll
            for (int j = 0; j < MMAX; ++j) {
                y0[j] = j < used ? Y[(size_t) j * A.ld + i] : (T) 0;
                y1[j] = (j < used && ok1) ? Y[(size_t) j * A.ld + i1] : (T) 0;
            }
            T t0 = -g_s[i - e0], t1 = ok1 ? -g_s[i1 - e0] : (T) 0;
            #pragma unroll
            for (int j = 0; j < MMAX; ++j) {
                if (j < used) { const T uj = (T) (-coef_s[m + j]); t0 = fma(uj, y0[j], t0); t1 = fma(uj, y1[j], t1); }
            }
            const double ht0 = (double) (h_s[i - e0] * t0);
            const double ht1 = ok1 ? (double) (h_s[i1 - e0] * t1) : 0.0;
            a_tt = fma(ht0, ht0, a_tt);
            a_tt = fma(ht1, ht1, a_tt);
            #pragma unroll
            for (int j = 0; j < MMAX; ++j) {
                if (j < used) { acc[j] = fma((double) y0[j], ht0, acc[j]); acc[j] = fma((double) y1[j], ht1, acc[j]); }
            }
        }
        {
            constexpr int NG = MMAX + 1 <= 16 ? 16 : 32;
            double v[NG];
            #pragma unroll
            for (int j = 0; j < NG; ++j) v[j] = j < MMAX ? acc[j < MMAX ? j : 0] : j == MMAX ? a_tt : 0.0;
            const double r = warp_reduce_multi<NG>(v);
            constexpr int SH = NG == 16 ? 1 : 0;              // lane l holds value l >> SH
            if ((lane & ((1 << SH) - 1)) == 0 && (lane >> SH) <= MMAX) red[warp][lane >> SH] = r;
        }
        __syncthreads();
        for (int p = tid; p < P2; p += kAdaThreads) {        // record of ka2_wdots: [0,m) w, [m] sum (h.(Yu-g))^2
            const int q = p < m ? p : MMAX;
            double v = 0;
            if (p == m || p < used) { for (int w2 = 0; w2 < NW; ++w2) v += red[w2][q]; }
            rec2[(size_t) p * G + b] = v;
        }
        stamp(5);
        barrier();
        stamp(6);

        // ---- C: a = R^-T (D u + w), bound, decision ----
        reduce_entry_major(rec2, P2, G, sums2, NW);
        for (int i = tid; i < used; i += kAdaThreads) sh.w[i] = sh.Rm[i][i] * sh.u[i] + sums2[ph(i)];
        __syncthreads();
        if (tid < 32) {
            const unsigned full = 0xffffffffu;
            const bool live = tid < used;
            const double tt = sums2[m];
            double U = sqrt(tt);
            bool ok = isfinite(tt);
            const double inv = 1.0 / (live ? sh.Rm[tid][tid] : 1.0);
            double t = live ? sh.w[tid] : 0.0, ai = 0.0;
            for (int j = 0; j < used; ++j) {
                const double aj = __shfl_sync(full, t * inv, j);
                if (tid == j) ai = aj;
                if (live && tid > j) t = fma(-sh.Rm[j][tid], aj, t);
            }
            double term = 0.0;
            bool fin = true;
            if (live) {
                coef_s[ph(tid)] = ai;
                term = fabs(ai) * sqrt(sh.ssv[tid]);
                fin = isfinite(ai) && isfinite(sh.u[tid]);
            }
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) term += __shfl_xor_sync(full, term, o);
            U += term;
            ok = ok && __all_sync(full, fin) && isfinite(U);
            if (tid == 0) {
                int s2 = ST_ACCEPT;
                if (A.check_nan) {
                    if (!ok) s2 = ST_REJECT_NONFINITE;
                    else if (!(U <= 0.99 * A.limit)) s2 = ST_NEED_EXACT_NORM;
                }
                sh.status = s2;
            }
        }
        __syncthreads();
        status = sh.status;
    }
    stamp(7);

    // d on the slice (ka3_combine): part0 = sum_r a_r s_r ; part1 = h.(g + sum_r b_r y_r)
    T cfS[MMAX], cfY[MMAX];
    #pragma unroll
    for (int r = 0; r < MMAX; ++r) { cfS[r] = r < used ? (T) coef_s[r] : (T) 0; cfY[r] = r < used ? (T) coef_s[m + r] : (T) 0; }
    auto direction = [&](long long i) -> T {
        T sv[MMAX], yv[MMAX];
        #pragma unroll
        for (int r = 0; r < MMAX; ++r) { sv[r] = r < used ? S[(size_t) r * A.ld + i] : (T) 0; yv[r] = r < used ? Y[(size_t) r * A.ld + i] : (T) 0; }
        const T gv = g_s[i - e0], h = h_s[i - e0];
        T p0 = (T) 0, p1 = gv;
        #pragma unroll
        for (int r = 0; r < MMAX; ++r) {
            if (r < used) { p0 = fma(cfS[r], sv[r], p0); p1 = fma(cfY[r], yv[r], p1); }
        }
        p1 = used > 0 ? h * p1 : h;
        return p1 + p0;
    };

    bool d_in_g = false;
    if (status == ST_NEED_EXACT_NORM) {                      // measure ||d|| exactly before touching x (stochqn.c:825-838)
        double a_dd = 0, a_bad = 0;
        T* gw = const_cast<T*>(g);
        for (long long i = e0 + tid; i < e1; i += kAdaThreads) {
            const T d = direction(i);
            const double de = (double) d;
            a_dd = fma(de, de, a_dd);
            if (!isfinite(de)) a_bad += 1.0;
            gw[i] = d;
        }
        a_dd = warp_sum(a_dd); a_bad = warp_sum(a_bad);
        __syncthreads();
        if (lane == 0) { red[warp][0] = a_dd; red[warp][1] = a_bad; }
        __syncthreads();
        if (tid < 2) { double v = 0; for (int w2 = 0; w2 < NW; ++w2) v += red[w2][tid]; partials[(size_t) tid * G + b] = v; }
        barrier();
        reduce_entry_major(partials, 2, G, two_s, NW);
        status = (two_s[1] > 0 || !(sqrt(two_s[0]) <= A.limit)) ? ST_REJECT_NONFINITE : ST_ACCEPT;
        d_in_g = true;
    }

    const T nstep = -step;
    if (status == ST_ACCEPT) {
        for (long long i = e0 + tid; i < e1; i += kAdaThreads) {
            const T d = d_in_g ? g[i] : direction(i);
            const T xv = fma(nstep, d, x[i]);
            x[i] = xv;
            x_sum[i] = x_sum[i] + xv;
            if (gout) gout[i] = d;
        }
    } else {
        for (long long i = e0 + tid; i < e1; i += kAdaThreads) x_sum[i] = x_sum[i] + x[i];      // quirk Q7 (stochqn.c:1191)
    }

    if (trace) { __syncthreads(); stamp(8); }
    R.last_status = status;
    R.calls += 1;
    if (A.fisher_size > 0) {                                 // add_to_fisher_mem (stochqn.c:581-587)
        R.f_st = (f_st + 1) % A.fisher_size;
        R.f_used = f_used + 1 >= A.fisher_size ? A.fisher_size : f_used + 1;
    }
    if (status == ST_ACCEPT) {
        R.pend = -1;
        R.n_ok += 1; R.x_changed += 1; R.last_info = 200;
    } else {
        R.used = 0; R.slot = 0; R.pend = -1;                 // flush_bfgs_mem (stochqn.c:554-558)
        R.n_nan += 1; R.last_info = 203;
    }
}

template <typename T, int MMAX>
__global__ void __launch_bounds__(kAdaThreads, 1)
kl_ada(const AdaLoopArgs A, LoopState* __restrict__ st, const T* g, T* gout, T* __restrict__ Gacc,
       const T* __restrict__ S, const T* __restrict__ Y, T* __restrict__ F, T* __restrict__ x, T* __restrict__ x_sum, const T step,
       double* partials, double* rec2, double* SY, double* YY, double* SS, unsigned long long* bar, unsigned long long* trace)
{
    extern __shared__ __align__(16) unsigned char ada_dyn_smem[];
    AdaRing R = ada_ring_load(st);
    kl_ada_body<T, MMAX>(A, R, g, gout, Gacc, S, Y, F, x, x_sum, step, partials, rec2, SY, YY, SS, [&]() { fit_barrier(bar); }, trace, ada_dyn_smem);
    // (every CTA passed at least one grid barrier after reading the record)
    if (blockIdx.x == 0 && threadIdx.x == 0) ada_ring_store(st, R);
}

// ---- adaQN + multinomial model: a RUN of ordinary mini-batches in one persistent launch --------------------------------------
// gradient of mini-batch ib (mn_grad_small_body, mn_small.cuh) -> barrier -> adaQN step (kl_ada_body) -> barrier -> next.
// The two bodies use the dynamic shared memory one after the other.  stochqn/_optimizers.py:339-382 for the common path.
template <typename T>
struct MnFitArgs {
    const T* X; long long ldx;          // first row of the first mini-batch
    const T* Y; long long ldy;          // one-hot labels
    const T* sw;                        // or nullptr
    long long batch_rows, rows_total;
    int nbatches, d, K, icpt;
    T alpha;
    T* Zp; T* Dg;                       // scratch of mn_grad_small
};

template <typename T, int MMAX>
__global__ void __launch_bounds__(kAdaThreads, 1)
kl_fit_mn_ada(const MnFitArgs<T> F, const AdaLoopArgs A, LoopState* __restrict__ st, T* g, T* __restrict__ Gacc,
              const T* __restrict__ S, const T* __restrict__ Y, T* __restrict__ Fm, T* x, T* __restrict__ x_sum, const T step,
              double* partials, double* rec2, double* SY, double* YY, double* SS, unsigned long long* bar, unsigned long long* trace)
{
    extern __shared__ __align__(16) unsigned char fit_dyn_smem[];
    AdaRing R = ada_ring_load(st);
    RunBarrier barrier{bar, 0ull};                           // (the host zeroes the counter before every launch of this kernel)
    for (int ib = 0; ib < F.nbatches; ++ib) {
        const long long r0 = (long long) ib * F.batch_rows;
        const long long rows = F.rows_total - r0 < F.batch_rows ? F.rows_total - r0 : F.batch_rows;
        mnsmall::MnSmallArgs<T> ma;
        ma.X = F.X + (size_t) r0 * (size_t) F.ldx; ma.ldx = F.ldx;
        ma.Y = F.Y + (size_t) r0 * (size_t) F.ldy; ma.ldy = F.ldy;
        ma.labels = nullptr;
        ma.sw = F.sw ? F.sw + r0 : nullptr;
        ma.B = (int) rows; ma.d = F.d; ma.K = F.K; ma.icpt = F.icpt;
        ma.W = x; ma.alpha = F.alpha; ma.Gout = g; ma.Zp = F.Zp; ma.Dg = F.Dg;
        mnsmall::mn_grad_small_body<T>(ma, fit_dyn_smem, barrier, nullptr);
        barrier();                                           // the gradient is complete
        kl_ada_body<T, MMAX>(A, R, g, (T*) nullptr, Gacc, S, Y, Fm, x, x_sum, step, partials, rec2, SY, YY, SS, barrier,
                             ib == F.nbatches - 1 ? trace : nullptr, fit_dyn_smem);
        barrier();                                           // x is complete
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) ada_ring_store(st, R);
}

}  // namespace sqn
