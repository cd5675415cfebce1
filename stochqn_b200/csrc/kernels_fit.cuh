// kernels_fit.cuh - a RUN of consecutive mini-batches of the guided request loop in ONE launch (binary logistic models).
//
// kernels_loop.cuh moved the decisions of a mini-batch to the device but still enqueues 4-6 kernels per mini-batch
// (gradient sweep + finishing pass, step, gradient sweep + finishing pass, pair): at the reference's own sizes
// (n = 1001 ... 4097, 8 KB - 32 KB vectors, batches of 1000-2000 rows) the chain of launch gaps and half-empty
// kernels is all there is to the step time.  kl_fit_logistic is a persistent cooperative grid - one 512-thread CTA per
// SM - that runs `nbatches` mini-batches back to back:
//
//   stochqn/_optimizers.py:339-382 (`_fit_batch`) / R/optimizers_guided.R:26-111     the request loop
//   R/logistic.R:12-21 or scikit-learn's _logistic_loss_and_grad (logistic_form.cuh)  the gradient that serves it
//   src/stochqn.c:802-840, 663-708, 915-926, 883-900                                 take_step, the two-loop, the pair, curvature
//
// per mini-batch (oLBFGS; SQN's ordinary steps stop after S):
//   R   rows     CTA b sweeps its share of the batch rows once (the one-sweep form of callbacks.cu: a CTA-wide column
//                accumulator in registers, R rows in flight, z_i = x_i'w reduced through shared memory, r_i * x_i added from
//                the registers that still hold the row) and leaves one partial column record              -- grid barrier
//   C   columns  every CTA owns a fixed slice of the n elements (<= 62): it adds its slice of the records in CTA order,
//                finishes the gradient there (mean / penalty / intercept), keeps it in shared memory, and forms its part
//                of the 4m+2 inner products of K1 (kernels.cuh) with its slices of S and Y                -- grid barrier
//   S   step     every CTA sums the partial records in the same order, solves the m x m compact-form system redundantly
//                (solve_cta), takes the accept / reject decision (exact-norm route: + one barrier), and updates ITS slice
//                of x, the new s (oLBFGS) or x_sum (SQN)                                                 -- grid barrier
//   R', C'       the gradient again at the new x on the same rows (now L2 hits), y = g' - g (+ y_reg s) on the slice,
//                partial s'y, s's                                                                       -- 2 grid barriers
//   P   pair     every CTA sums the 2-value records, applies check_min_curvature and quirk Q1 to its slice, advances
//                its copy of the ring counters
//
// The ring counters live in registers of every CTA (all CTAs take identical decisions from identical sums); CTA 0
// writes them to the LoopState record once, at the end.  The pair decision P rides on the first barrier of the NEXT
// mini-batch (whose row sweep does not depend on it): four grid barriers per oLBFGS iteration, none of them a launch.
// Data written by one CTA and read by another inside the launch (x, the records, the Gram state) is read with
// ld.global.cg: L1 is not coherent across SMs.  S / Y / x_sum slices are only ever touched by their owner.
#pragma once

#include "logistic_form.cuh"

namespace sqn {

constexpr int kFitThreads = 512;
constexpr int kFitWarps = kFitThreads / 32;
constexpr int kFitMaxPer = 62;                  // elements of an n-vector a CTA may own (+ 2 pseudo-columns = 64 lanes of work)
constexpr int kFitSlices = kFitThreads / 64;    // record slices of the column phase

template <typename T>
struct FitArgs {
    const T* X;                 // first row of the first mini-batch
    const T* y;
    const T* sw;                // or nullptr
    long long ldx, ncols;       // ncols: columns of X (the sk form with an intercept has n = ncols + 1 variables)
    long long batch_rows;       // rows per mini-batch
    long long rows_total;       // rows available from X on (the last mini-batch may be cut short)
    int nbatches;
    int sk, icpt;
    double lambda;
    unsigned long long* trace;  // development aid: CTA 0 stamps %globaltimer at the phase boundaries of every mini-batch (16 slots each), or nullptr
};

template <typename T, int MODE, int CPT, int R>
__global__ void __launch_bounds__(kFitThreads, 1)
kl_fit_logistic(const FitArgs<T> F, const LoopArgs A, LoopState* __restrict__ st, T* x, T* x_sum, T* S, T* Y, const T step,
                double* colpart, double* partials, double* rec2, double* SY, double* YY, double* SS,
                unsigned long long* bar)
{
    extern __shared__ __align__(16) unsigned char fit_smem[];
    T* ws = reinterpret_cast<T*>(fit_smem);                    // the point of evaluation, ws[c] = x[c], c < CPT * 512
    __shared__ double red[2][kFitWarps][R];
    __shared__ double rw_s[2][2 * R];
    __shared__ double part[kFitSlices][64];
    __shared__ double sums_s[4 * kMaxMem + 2];
    __shared__ double coef_s[2 * kMaxMem + 3];
    __shared__ double two_s[2];
    __shared__ double tot_s[64];
    __shared__ SolveShared sh;
    __shared__ T g_s[64], gp_s[64], s_s[64];

    const int m = A.msize, P = 4 * m + 2;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = (int) gridDim.x, b = (int) blockIdx.x;
    const long long n = A.n, ncols = F.ncols;
    const long long per = (n + G - 1) / G;
    const long long e0 = (long long) b * per < n ? (long long) b * per : n;
    const long long e1 = e0 + per < n ? e0 + per : n;
    const int ne = (int) (e1 - e0);
    const size_t rec = (size_t) (ncols + 2);
    const LgForm form{F.sk, F.icpt};
    const T nstep = -step;
    const T y_reg = (T) A.y_reg;

    RunBarrier barrier{bar, 0ull};                             // (the host zeroes the counter before every launch of this kernel)
    int used = st->used, slot = st->st_ix, pend = st->pend;
    int last_info = st->last_info, last_status = st->last_status;
    unsigned long long n_ok = 0, n_curv = 0, n_nan = 0, calls = 0, x_changed = 0;

    // ---- R: one sweep of the rows [r0, r0 + rows) of the batch at the point ws -> this CTA's partial column record ----
    auto row_phase = [&](const T* Xb, const T* yb, const T* swb, long long rows) {
        // the point: every thread keeps the entries of its own columns (registers when there are few, else its own
        // shared-memory slots); the loads are issued here and first used after the row loads of the first chunk
        constexpr bool WREG = CPT <= 4;
        T wk[CPT];
        #pragma unroll
        for (int k = 0; k < CPT; ++k) {
            const long long c = tid + (long long) k * kFitThreads;
            wk[k] = c < n ? __ldcg(x + c) : (T) 0;
        }
        const double zc = form.icpt ? (double) __ldcg(x + ncols) : 0.0;
        if (!WREG) {
            #pragma unroll
            for (int k = 0; k < CPT; ++k) ws[tid + k * kFitThreads] = wk[k];
        }
        T acc[CPT];
        #pragma unroll
        for (int k = 0; k < CPT; ++k) acc[k] = (T) 0;
        double sw_sum = 0.0, r_sum = 0.0;
        const long long rpc = (rows + G - 1) / G;
        const long long rb = (long long) b * rpc;
        const long long re = rb + rpc < rows ? rb + rpc : rows;
        int par = 0;
        bool first = WREG;
        for (long long row0 = rb; row0 < re; row0 += R) {
            T xv[R][CPT];
            #pragma unroll
            for (int r = 0; r < R; ++r) {
                const long long row = row0 + r;
                const T* xr = Xb + row * F.ldx;
                #pragma unroll
                for (int k = 0; k < CPT; ++k) {
                    const long long c = tid + (long long) k * kFitThreads;
                    xv[r][k] = (row < re && c < ncols) ? __ldg(xr + c) : (T) 0;
                }
            }
            if (first) {
                #pragma unroll
                for (int k = 0; k < CPT; ++k) ws[tid + k * kFitThreads] = wk[k];
                first = false;
            }
            T zp[R];
            #pragma unroll
            for (int r = 0; r < R; ++r) zp[r] = (T) 0;
            #pragma unroll
            for (int k = 0; k < CPT; ++k) {
                const long long c = tid + (long long) k * kFitThreads;
                T wv = WREG ? wk[k] : ws[tid + k * kFitThreads];
                if (c >= ncols) wv = (T) 0;                                  // (the sk intercept is not a column of X)
                #pragma unroll
                for (int r = 0; r < R; ++r) zp[r] = fma(xv[r][k], wv, zp[r]);
            }
            #pragma unroll
            for (int r = 0; r < R; ++r) {
                const double z = warp_sum((double) zp[r]);
                if (lane == 0) red[par][warp][r] = z;
            }
            __syncthreads();
            // the row weights once per row (thread r < R), not once per thread: exp + divide in fp64 are ~100 issue slots,
            // 512 threads doing all R of them was half of the sweep time at 1000-column rows
            if (tid < R) {
                const long long row = row0 + tid;
                double z = 0;
                #pragma unroll
                for (int q = 0; q < kFitWarps; ++q) z += red[par][q][tid];
                double rr = 0.0, wt = 0.0;
                if (row < re) {
                    wt = swb ? (double) swb[row] : 1.0;
                    rr = lg_row_weight(LG_GRAD, form, z + zc, 0.0, (double) yb[row], wt);
                }
                rw_s[par][tid] = rr;
                rw_s[par][R + tid] = wt;
            }
            __syncthreads();
            #pragma unroll
            for (int r = 0; r < R; ++r) {
                const double rr = rw_s[par][r];
                if (tid == 0) { sw_sum += rw_s[par][R + r]; r_sum += rr; }
                const T rt = (T) rr;
                #pragma unroll
                for (int k = 0; k < CPT; ++k) acc[k] = fma(rt, xv[r][k], acc[k]);
            }
            par ^= 1;
        }
        if (first) {                                                         // (a CTA without rows still owes the point to its column phase)
            #pragma unroll
            for (int k = 0; k < CPT; ++k) ws[tid + k * kFitThreads] = wk[k];
        }
        double* out = colpart + (size_t) b * rec;
        #pragma unroll
        for (int k = 0; k < CPT; ++k) {
            const long long c = tid + (long long) k * kFitThreads;
            if (c < ncols) out[c] = (double) acc[k];
        }
        if (tid == 0) { out[ncols] = sw_sum; out[ncols + 1] = r_sum; }
    };

    // ---- C: gradient on the element slice [e0, e1) from the records of all CTAs -> g_s (fixed order) ------------------
    auto col_phase = [&]() {
        const int j = tid & 63, q = tid >> 6;
        // pseudo-columns ne, ne + 1: the sums of the sample weights and of the row weights
        const long long col = j < ne ? e0 + j : (j == ne ? ncols : ncols + 1);
        const bool have = j < ne + 2 && !(j < ne && col >= ncols);           // (col == ncols inside the slice: the sk intercept, taken from r_sum)
        double asum = 0;
        if (have) {
            // this thread's records q, q + 8, ...: up to ten loads in flight per trip (two trips cover 160 CTAs), added in record order
            const double* pc = colpart + col;
            for (int r0 = q; r0 < G; r0 += 10 * kFitSlices) {
                double v[10];
                #pragma unroll
                for (int u = 0; u < 10; ++u) { const int r = r0 + u * kFitSlices; v[u] = r < G ? __ldcg(pc + (size_t) r * rec) : 0.0; }
                #pragma unroll
                for (int u = 0; u < 10; ++u) asum += v[u];
            }
        }
        part[q][j] = asum;
        __syncthreads();
        if (tid < 64) {
            double t = 0;
            #pragma unroll
            for (int s = 0; s < kFitSlices; ++s) t += part[s][tid];
            tot_s[tid] = t;
        }
        __syncthreads();
        if (tid < ne) {
            const long long e = e0 + tid;
            const double swt = tot_s[ne], rst = tot_s[ne + 1];
            double g;
            if (e >= ncols) g = rst;                                           // unpenalised intercept of the sk form
            else g = form.sk ? tot_s[tid] + F.lambda * (double) ws[e] : tot_s[tid] / swt + 2.0 * F.lambda * (double) ws[e];
            g_s[tid] = (T) g;
        }
        __syncthreads();
    };

    auto stamp = [&](int ib, int k) {
        if (F.trace && b == 0 && tid == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            F.trace[(size_t) ib * 16 + k] = t;
        }
    };

    // ---- P: check_min_curvature (stochqn.c:883-900) on the sums of the 2-value records, quirk Q1 on rejection, ring advance.
    // Every CTA takes the same decision from the same sums; call it after a grid barrier that follows the records.
    bool pair_pending = false;
    auto pair_decision = [&]() {
        if (warp < 2) {
            double v = 0;
            for (int r0 = 0; r0 < G; r0 += 5 * 32) {
                double t[5];
                #pragma unroll
                for (int u = 0; u < 5; ++u) { const int r = r0 + u * 32 + lane; t[u] = r < G ? __ldcg(rec2 + (size_t) r * 2 + warp) : 0.0; }
                #pragma unroll
                for (int u = 0; u < 5; ++u) v += t[u];
            }
            v = warp_sum(v);
            if (lane == 0) two_s[warp] = v;
        }
        __syncthreads();
        const bool reject = A.min_curvature > 0 && ((T) two_s[0] / (T) two_s[1]) <= (T) A.min_curvature;      // the division in T (stochqn.c:892)
        calls += 1;
        if (reject) {                                             // quirk Q1: the slot is zeroed, the counters stay
            if (tid < ne) { S[(size_t) slot * A.ld + e0 + tid] = (T) 0; Y[(size_t) slot * A.ld + e0 + tid] = (T) 0; }
            if (b == 0) {
                for (int j = tid; j < m; j += kFitThreads) {
                    SY[j * m + slot] = 0; SY[slot * m + j] = 0;
                    YY[j * m + slot] = 0; YY[slot * m + j] = 0;
                }
                if (tid == 0) SS[slot] = 0;
            }
            n_curv += 1; last_info = 202;
        } else {
            pend = slot;
            slot = (slot + 1) % m;                                // incr_bfgs_counters (stochqn.c:569-573)
            used = used + 1 >= m ? m : used + 1;
            n_ok += 1; last_info = 200;
        }
        pair_pending = false;
        __syncthreads();                                          // two_s is reused
    };

    for (int ib = 0; ib < F.nbatches; ++ib) {
        stamp(ib, 0);
        const long long r0 = (long long) ib * F.batch_rows;
        const long long rows = F.rows_total - r0 < F.batch_rows ? F.rows_total - r0 : F.batch_rows;
        const T* Xb = F.X + (size_t) r0 * (size_t) F.ldx;
        const T* yb = F.y + r0;
        const T* swb = F.sw ? F.sw + r0 : nullptr;

        row_phase(Xb, yb, swb, rows);
        stamp(ib, 1);
        barrier();
        stamp(ib, 2);
        if (pair_pending) pair_decision();                        // the previous mini-batch's pair
        col_phase();
        stamp(ib, 3);

        // ---- partial inner products of K1 on the slice (kernels.cuh: record layout of k1_dots) ----
        {
            double* prec = partials + b;                        // entry-major records: entry p of CTA b is partials[p * G + b]
            for (int p = tid; p < P; p += kFitThreads) prec[(size_t) p * G] = 0.0;
            __syncthreads();
            const int c = pend;
            const int nd = 2 * used + 1 + (c >= 0 ? 2 * used + 1 : 0);
            constexpr int DPW = 3;                             // dots per warp and round, all their loads issued before the first use
            for (int d0 = 0; d0 < nd; d0 += DPW * kFitWarps) {
                T av[DPW][2], bv[DPW][2];
                int idx[DPW];
                #pragma unroll
                for (int u = 0; u < DPW; ++u) {
                    const int d = d0 + warp + u * kFitWarps;
                    idx[u] = -1;
                    av[u][0] = av[u][1] = bv[u][0] = bv[u][1] = (T) 0;
                    if (d < nd) {
                        const T* a;
                        const T* bb = nullptr;                                          // nullptr: the gradient (shared memory)
                        if (d < used)              { a = S + (size_t) d * A.ld;                  idx[u] = d; }
                        else if (d < 2 * used)     { a = Y + (size_t) (d - used) * A.ld;         idx[u] = m + (d - used); }
                        else if (d == 2 * used)    { a = nullptr;                                idx[u] = 4 * m; }
                        else if (d < 3 * used + 1) { a = S + (size_t) (d - 2 * used - 1) * A.ld; bb = Y + (size_t) c * A.ld; idx[u] = 2 * m + (d - 2 * used - 1); }
                        else if (d < 4 * used + 1) { a = Y + (size_t) (d - 3 * used - 1) * A.ld; bb = Y + (size_t) c * A.ld; idx[u] = 3 * m + (d - 3 * used - 1); }
                        else                       { a = S + (size_t) c * A.ld;                  bb = a; idx[u] = 4 * m + 1; }
                        #pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int jj = lane + 32 * h;
                            if (jj < ne) {
                                av[u][h] = a ? a[e0 + jj] : g_s[jj];
                                bv[u][h] = bb ? bb[e0 + jj] : g_s[jj];
                            }
                        }
                    }
                }
                #pragma unroll
                for (int u = 0; u < DPW; ++u) {
                    if (idx[u] >= 0) {                                                  // warp-uniform
                        double v = fma((double) av[u][0], (double) bv[u][0], 0.0);
                        v = fma((double) av[u][1], (double) bv[u][1], v);
                        v = warp_sum(v);
                        if (lane == 0) prec[(size_t) idx[u] * G] = v;
                    }
                }
            }
            if (MODE == MODE_OLBFGS && tid < ne) gp_s[tid] = g_s[tid];                  // grad_prev <- grad (stochqn.c:996)
        }
        stamp(ib, 4);
        barrier();
        stamp(ib, 5);

        // ---- S: every CTA reduces the records in the same order, solves, decides, updates its slice ----
        reduce_entry_major(partials, P, G, sums_s, kFitWarps);     // (entry-major records: lanes read consecutive CTAs' values)
        stamp(ib, 12);
        SolveArgs SA;
        SA.msize = m; SA.used = used; SA.oldest = (slot == used) ? 0 : slot; SA.pend = pend; SA.nblocks = 0; SA.do_solve = 1;
        SA.check_nan = A.check_nan; SA.h0 = A.h0; SA.limit = A.limit; SA.seq = 0;
        int status = solve_cta(SA, sums_s, SY, YY, SS, sh, coef_s, b == 0, true, kFitThreads);
        stamp(ib, 13);
        const T gamma = (T) coef_s[2 * m];
        T d = (T) 0;
        if (tid < ne) d = combine_direction_at<T>(g_s[tid], S, Y, A.ld, e0 + tid, used, m, gamma, coef_s);
        stamp(ib, 14);
        if (status == ST_NEED_EXACT_NORM) {                      // the reference's check-before-update order (stochqn.c:825-838)
            double a_dd = 0, a_bad = 0;
            if (tid < ne) { const double de = (double) d; a_dd = de * de; if (!isfinite(de)) a_bad = 1.0; }
            if (tid < 64) {
                a_dd = warp_sum(a_dd); a_bad = warp_sum(a_bad);
                if (lane == 0) { part[0][warp] = a_dd; part[1][warp] = a_bad; }
            }
            __syncthreads();
            if (tid == 0) { rec2[(size_t) b * 2] = part[0][0] + part[0][1]; rec2[(size_t) b * 2 + 1] = part[1][0] + part[1][1]; }
            barrier();
            if (warp < 2) {
                double v = 0;
                for (int r = lane; r < G; r += 32) v += __ldcg(rec2 + (size_t) r * 2 + warp);
                v = warp_sum(v);
                if (lane == 0) two_s[warp] = v;
            }
            __syncthreads();
            status = (two_s[1] > 0 || !(sqrt(two_s[0]) <= A.limit)) ? ST_REJECT_NONFINITE : ST_ACCEPT;
            __syncthreads();
        }
        calls += 1;
        last_status = status;
        if (status == ST_ACCEPT) {
            if (tid < ne) {
                const long long e = e0 + tid;
                const T xv = fma(nstep, d, ws[e]);
                x[e] = xv;
                if constexpr (MODE == MODE_OLBFGS) {
                    const T sv = nstep * d;
                    S[(size_t) slot * A.ld + e] = sv;
                    s_s[tid] = sv;
                } else {
                    x_sum[e] = x_sum[e] + xv;
                }
            }
            pend = -1;
            n_ok += 1; x_changed += 1; last_info = 200;
        } else {
            if constexpr (MODE == MODE_AVG) {
                if (tid < ne) x_sum[e0 + tid] = x_sum[e0 + tid] + ws[e0 + tid];          // quirk Q7 (stochqn.c:1067)
            }
            used = 0; slot = 0; pend = -1;                        // flush_bfgs_mem (stochqn.c:554-558)
            n_nan += 1; last_info = 203;
        }
        stamp(ib, 6);
        barrier();                                  // x is complete
        stamp(ib, 7);

        if (MODE != MODE_OLBFGS || status != ST_ACCEPT) continue;

        // ---- R', C', P: the pair (stochqn.c:915-926, 883-900) ----
        row_phase(Xb, yb, swb, rows);
        stamp(ib, 8);
        barrier();
        col_phase();
        stamp(ib, 9);
        {
            double a_sy = 0, a_ss = 0;
            T yv = (T) 0;
            if (tid < ne) {
                const T sv = s_s[tid];
                yv = g_s[tid] - gp_s[tid];
                if (y_reg > (T) 0) yv = fma(y_reg, sv, yv);
                Y[(size_t) slot * A.ld + e0 + tid] = yv;
                a_sy = (double) sv * (double) yv;
                a_ss = (double) sv * (double) sv;
            }
            if (tid < 64) {
                a_sy = warp_sum(a_sy); a_ss = warp_sum(a_ss);
                if (lane == 0) { part[0][warp] = a_sy; part[1][warp] = a_ss; }
            }
            __syncthreads();
            if (tid == 0) { rec2[(size_t) b * 2] = part[0][0] + part[0][1]; rec2[(size_t) b * 2 + 1] = part[1][0] + part[1][1]; }
            stamp(ib, 10);
            // The decision needs the sums over all CTAs, but nothing before the NEXT mini-batch's column phase depends on it (its row
            // sweep only reads x): it rides on that mini-batch's first barrier (pair_decision below) - four barriers per iteration.
            pair_pending = true;
        }
    }
    if (pair_pending) {                                           // the last pair of the run
        barrier();
        pair_decision();
    }

    if (b == 0 && tid == 0) {
        st->used = used; st->st_ix = slot; st->pend = pend;
        st->skip_pair = 0;
        st->last_status = last_status; st->last_info = last_info;
        st->n_info[0] += n_ok; st->n_info[2] += n_curv; st->n_info[3] += n_nan;
        st->calls += calls; st->x_changed += x_changed;
    }
}

}  // namespace sqn
