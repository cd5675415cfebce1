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
//             the two reduction phases of the compact form with a diagonal H0 (kernels_adaqn.cuh), update
//
// All three are cooperative grids of 256-thread CTAs (or one 1024-thread CTA for n <= 2048) in which every CTA owns a
// contiguous slice of the elements; partial records are summed by every CTA in the same fixed order, and every CTA
// solves the m x m system redundantly in its own shared memory, so no broadcast is needed after a grid barrier.
// Arithmetic of the update is that of K3 / KA3 (same FMA order), so the iterates agree with the host-driven routes to
// the last bit given the same coefficients.
#pragma once

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
    // check_min_curvature (stochqn.c:883-900), in the precision the host route uses
    const bool reject = A.min_curvature > 0 && (two_s[0] / two_s[1]) <= A.min_curvature;
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

}  // namespace sqn
