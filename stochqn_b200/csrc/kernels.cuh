// kernels.cuh - device kernels of the B200-native stochastic quasi-Newton step (sm_100a).
//
// The reference computes H*g with the latency-chained two-loop recursion
// (src/stochqn.c:663-708): 4m dependent dot products and 2m axpys, each a full sweep
// over an n-vector (about 12m+5 vector transfers).  Here the same product is evaluated in
// the compact (Byrd-Nocedal-Schnabel) form, which needs only TWO sweeps over the
// correction pairs:
//
//   K1  k1_dots      one coalesced pass over g, S, Y: all S'g, Y'g (+ the Gram column of
//                    the newest pair, + g'g), fp64 accumulation, deterministic block partials
//   K2  k2_solve     one CTA: sums the partials, folds the new Gram column in, solves the
//                    m x m triangular systems, emits 2m+1 coefficients and the accept flag
//   K3  k3_combine   one coalesced pass: d = gamma*g + S a + gamma*Y b, x -= step*d, and the
//                    optimizer-specific epilogue (oLBFGS: s slot; SQN/adaQN: x_sum += x)
//   K4  k4_pair      y = g_new - g_prev (+ y_reg*s)  | y = hess_vec, with s'y and s's
//
// Everything is HBM-bound streaming work (about 0.25 flop/byte); there is no GEMM here and
// none is manufactured.  All kernels are templated on the storage type T (double / float);
// every reduction accumulates in fp64 whatever T is.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <type_traits>
#include "vecio.cuh"
#include "p2p.cuh"

namespace sqn {

constexpr int kThreads = 256;          // threads per CTA for the streaming kernels
constexpr int kWarps = kThreads / 32;
constexpr int kMaxMem = 32;            // largest mem_size handled by the compact solve

// status word written by K2 (device + mapped host copy)
enum : int { ST_ACCEPT = 0, ST_REJECT_NONFINITE = 1, ST_NEED_EXACT_NORM = 2, ST_COMM_TIMEOUT = 3 };

__device__ __forceinline__ double warp_sum(double v)
{
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic CTA reduction of `count` per-thread accumulators: lane 0 of every warp
// parks its warp sum in shared memory, then thread p adds the kWarps values in fixed order.
// `emit(p, value)` is called once per accumulator by the thread that owns it.
template <int COUNT_MAX, int THREADS = kThreads, typename Get, typename Emit>
__device__ __forceinline__ void block_reduce(int count, Get get, Emit emit)
{
    constexpr int NW = THREADS / 32;
    __shared__ double red[NW][COUNT_MAX];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    #pragma unroll
    for (int p = 0; p < COUNT_MAX; ++p) {
        if (p < count) {
            double v = warp_sum(get(p));
            if (lane == 0) red[warp][p] = v;
        }
    }
    __syncthreads();
    for (int p = threadIdx.x; p < count; p += THREADS) {
        double v = 0;
        #pragma unroll
        for (int w = 0; w < NW; ++w) v += red[w][p];
        emit(p, v);
    }
}

// =========================================================================================
// K1: fused multi-dot pass.
//
// Layout of one partial / sum record (m = mem_size, all fp64), index k*m + j for slot j:
//   k=0: s_j'g   k=1: y_j'g   k=2: s_j'y_c   k=3: y_j'y_c      (c = pending newest pair, rows sc_row / yc_row)
//   [4m] g'g     [4m+1] s_c's_c
// Rows are addressed by PHYSICAL slot; valid slots are always 0..used-1 (the ring fills
// from slot 0 after every flush and all slots are valid once it is full).
// Also writes grad_prev <- g when asked (the oLBFGS copy of stochqn.c:996 rides along).
// Replaces: stochqn.c:671-679 and 702-707 (the 4m chained dots), 996.
//
// Work split: the CTA is kGroups row-groups x kLanes chunk-lanes.  The 2*cnt "virtual rows"
// of a launch (S rows j0..j0+cnt-1, then Y rows j0..j0+cnt-1) are dealt RPG per group, so a
// thread streams 2-3 probe chunks + RPG row chunks and keeps only 2*RPG+2 fp64 accumulators:
// few registers, 3-4 CTAs per SM, every load of an iteration in flight at once.  The probe
// chunks (g, y_c) are fetched by all groups of the CTA - one DRAM read, then L1/L2 hits.
// =========================================================================================
constexpr int kGroups = 4;
constexpr int kLanes = kThreads / kGroups;     // 64: two warps per row-group

constexpr int kUnroll = 2;      // chunks per thread in flight (tools/kbench.cu: 2 x 8 loads x 512 threads/SM saturates HBM)

template <typename T, int RPG, bool PENDING, int VEC>
__global__ void __launch_bounds__(kThreads, 2)
k1_dots(const T* __restrict__ g, const T* __restrict__ S, const T* __restrict__ Y, size_t ld,
        int msize, int used, int j0, const T* __restrict__ sc_row, const T* __restrict__ yc_row, long long n,
        T* __restrict__ grad_prev, double* __restrict__ partials)
{
    // slots handled by this launch: j0 .. j0+cnt-1 (mem_size > 2*RPG takes several launches;
    // g'g, s_c's_c and the grad_prev copy belong to the launch with j0 == 0)
    int cnt = used - j0;
    if (cnt > 2 * RPG) cnt = 2 * RPG;               // kGroups*RPG = 4*RPG virtual rows = 2*RPG slots
    if (cnt < 0) cnt = 0;
    const int group = threadIdx.x / kLanes, lane = threadIdx.x % kLanes;
    const bool lead = (group == 0) && (j0 == 0);    // owns g'g and s_c's_c
    if (j0 > 0 || group != kGroups - 1) grad_prev = nullptr;
    const int nv = 2 * cnt;
    // row base pointers of this thread's RPG virtual rows (clamped: duplicates are dropped at the end)
    const T* rows[RPG];
    #pragma unroll
    for (int r = 0; r < RPG; ++r) {
        int v = group * RPG + r;
        if (v >= nv) v = nv > 0 ? nv - 1 : 0;
        rows[r] = (v < cnt || nv == 0) ? S + (size_t) (j0 + v) * ld : Y + (size_t) (j0 + v - cnt) * ld;
    }
    double a_g[RPG], a_c[RPG];
    double a_gg = 0, a_ss = 0;
    #pragma unroll
    for (int r = 0; r < RPG; ++r) { a_g[r] = 0; a_c[r] = 0; }

    // U chunks per thread: all loads of the U chunks are issued before the first use
    auto many = [&](const long long (&c)[kUnroll], auto vtag) {
        constexpr int V = decltype(vtag)::value;
        Pack<T, V> gv[kUnroll], yc[kUnroll], sc[kUnroll], rv[kUnroll][RPG];
        #pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (c[u] >= 0) {
                const size_t off = (size_t) c[u] * V;
                gv[u] = ld_stream<T, V>(g + off);
                if constexpr (PENDING) {
                    yc[u] = ld_stream<T, V>(yc_row + off);
                    if (lead) sc[u] = ld_stream<T, V>(sc_row + off);
                }
                #pragma unroll
                for (int r = 0; r < RPG; ++r) rv[u][r] = ld_row<T, V>(rows[r] + off);
            }
        }
        #pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (c[u] >= 0) {
                const size_t off = (size_t) c[u] * V;
                if (grad_prev) st_vec<T, V>(grad_prev + off, gv[u]);
                if (lead) {
                    #pragma unroll
                    for (int e = 0; e < V; ++e) {
                        double ge = (double) gv[u].get(e);
                        a_gg = fma(ge, ge, a_gg);
                        if constexpr (PENDING) { double se = (double) sc[u].get(e); a_ss = fma(se, se, a_ss); }
                    }
                }
                #pragma unroll
                for (int r = 0; r < RPG; ++r) {
                    #pragma unroll
                    for (int e = 0; e < V; ++e) {
                        double re = (double) rv[u][r].get(e);
                        a_g[r] = fma(re, (double) gv[u].get(e), a_g[r]);
                        if constexpr (PENDING) a_c[r] = fma(re, (double) yc[u].get(e), a_c[r]);
                    }
                }
            }
        }
    };
    const long long nchunks = n / VEC;
    const long long stride = (long long) gridDim.x * kLanes;
    for (long long c0 = (long long) blockIdx.x * kLanes + lane; c0 < nchunks; c0 += stride * kUnroll) {
        long long c[kUnroll];
        #pragma unroll
        for (int u = 0; u < kUnroll; ++u) { c[u] = c0 + u * stride; if (c[u] >= nchunks) c[u] = -1; }
        many(c, std::integral_constant<int, VEC>{});
    }
    if (VEC > 1 && blockIdx.x == 0) {               // scalar tail: n not a multiple of VEC
        const long long i = nchunks * VEC + lane;
        long long c[kUnroll];
        #pragma unroll
        for (int u = 0; u < kUnroll; ++u) c[u] = -1;
        if (i < n) { c[0] = i; many(c, std::integral_constant<int, 1>{}); }
    }

    // ---- CTA reduction: two warps per group, fixed order ----
    constexpr int NA = 2 * RPG + 2;
    __shared__ double red[kWarps][NA];
    const int warp = threadIdx.x >> 5, wl = threadIdx.x & 31;
    #pragma unroll
    for (int p = 0; p < NA; ++p) {
        double v = p < RPG ? a_g[p < RPG ? p : 0] : p < 2 * RPG ? a_c[(p - RPG) < RPG ? (p - RPG) : 0] : p == 2 * RPG ? a_gg : a_ss;
        v = warp_sum(v);
        if (wl == 0) red[warp][p] = v;
    }
    __syncthreads();
    const int P = 4 * msize + 2;
    double* out = partials + (size_t) blockIdx.x * P;
    constexpr int WPG = kWarps / kGroups;            // warps per group
    for (int t = threadIdx.x; t < kGroups * NA; t += kThreads) {
        const int gi = t / NA, p = t % NA;
        double v = 0;
        #pragma unroll
        for (int w = 0; w < WPG; ++w) v += red[gi * WPG + w][p];
        if (p >= 2 * RPG) {                          // scalars: only the lead group's are meaningful
            if (gi == 0 && j0 == 0) out[4 * msize + (p - 2 * RPG)] = v;
            continue;
        }
        const int r = p % RPG, vrow = gi * RPG + r;
        if (vrow >= nv) continue;                    // clamped duplicate
        const bool is_s = vrow < cnt;
        const int j = j0 + (is_s ? vrow : vrow - cnt);
        const int k = (p < RPG) ? (is_s ? 0 : 1) : (is_s ? 2 : 3);
        if (k < 2 || PENDING) out[k * msize + j] = v;
    }
    // entries this launch owns but did not compute: slots >= used (first launch) and the k=2,3 blocks without a pending pair
    if (j0 == 0) {
        for (int t = threadIdx.x; t < 4 * msize; t += kThreads) {
            const int k = t / msize, j = t % msize;
            if (j >= used || (!PENDING && k >= 2)) out[t] = 0.0;
        }
    }
}

// =========================================================================================
// K2: partial sums -> Gram update -> compact-form coefficients (one CTA).
//
// Gram state (fp64, physical-slot indexed, row-major m x m):  SY[i][j] = s_i'y_j,
// YY[i][j] = y_i'y_j, SS[i] = s_i's_i.  With pairs in logical order oldest..newest,
// R = upper(SY), D = diag(SY), p = S'g, q0 = Y'g, and H0 = gamma*I:
//     u = R^-1 p ;  b = -u ;  a = R^-T [ (D + gamma*YY) u - gamma*q0 ]
//     H g = gamma*g + S a + gamma * Y b
// gamma = hess_init if > 0, else s_l'y_l / y_l'y_l of the newest pair (stochqn.c:683-699).
// With no pairs: d = g (stochqn.c:808-812).  The triangular solves propagate Inf/NaN exactly
// where the two-loop's rho = 1/(y's) would (a zeroed slot gives 0/0).
//
// Accept test (stochqn.c:825-835): the reference rejects when the direction has a non-finite
// entry or ||d|| > 1e3*n.  Here: non-finite sums / coefficients -> ST_REJECT_NONFINITE;
// otherwise the triangle-inequality bound U >= ||d|| (no cancellation) - if U <= 0.99*limit the
// step is certainly acceptable and K3 may update x in the same pass; else ST_NEED_EXACT_NORM
// makes the host take the two-pass route that measures ||d|| exactly before touching x.
//
// coef layout: [0,m) a by physical slot, [m,2m) gamma*b by physical slot, [2m] gamma,
//              [2m+1] U bound, [2m+2] g'g.
// =========================================================================================
struct SolveArgs {
    int msize, used, oldest, pend;      // pend < 0: no pending pair
    int nblocks;                        // partial records to sum (0: sums already reduced, e.g. after an all-reduce)
    int do_solve;                       // 0: only reduce partials into sums (multi-GPU phase A)
    int check_nan;
    double h0;                          // hess_init (oLBFGS) or 0
    double limit;                       // 1e3 * n_global
    unsigned long long seq;             // published to the host flag block after the status word (host polls it)
};

// what the host polls: payload first, then a system-scope fence, then the sequence number
__device__ __forceinline__ void publish_seq(volatile unsigned long long* seq_host, unsigned long long seq)
{
    __threadfence_system();
    *seq_host = seq;
}

__device__ __forceinline__ bool finite_d(double v) { return isfinite(v); }

// Sum the CTA partial records into `sums` in a fixed order: one warp per entry, lanes stride over CTAs (lane-sequential
// over its CTAs, then a shuffle tree - the order never depends on timing).  A warp works on up to EPW entries at once so
// that EPW independent loads are in flight per lane: the loop is latency-bound (L2 hits), and with one entry at a time
// it cost ~15 us for 296 records of 42 values; this form ~4 us.
template <int EPW = 6>
__device__ __forceinline__ void reduce_partials(const double* __restrict__ partials, int nblocks, int P, double* __restrict__ sums)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int p0 = 0; p0 < P; p0 += kWarps * EPW) {
        double acc[EPW];
        #pragma unroll
        for (int q = 0; q < EPW; ++q) acc[q] = 0.0;
        #pragma unroll 2
        for (int b = lane; b < nblocks; b += 32) {
            const double* rec = partials + (size_t) b * P + p0 + warp;
            #pragma unroll
            for (int q = 0; q < EPW; ++q) if (p0 + warp + q * kWarps < P) acc[q] += __ldcg(rec + q * kWarps);
        }
        #pragma unroll
        for (int q = 0; q < EPW; ++q) {
            const int p = p0 + warp + q * kWarps;
            if (p < P) {                                  // warp-uniform
                const double v = warp_sum(acc[q]);
                if (lane == 0) sums[p] = v;
            }
        }
    }
}

// The solve proper, shared by k2_solve (one CTA after K1) and ks_step (every CTA of the fused small-n step,
// kernels_small.cuh).  Whole-CTA call.  `sums` holds the reduced record.  The Gram state is read from global memory
// with the entries of a pending pair taken from `sums` (so a CTA never depends on another CTA's Gram write);
// `write_gram` makes this CTA fold the pending column into the global Gram state.  The coefficients land in
// `coef_s` (shared memory, 2m+3 doubles: same layout as `coef`); the return value is the status word, valid in
// every thread.
struct SolveShared {
    double Rm[kMaxMem][kMaxMem + 1], Yl[kMaxMem][kMaxMem + 1];
    double pv[kMaxMem], qv[kMaxMem], ssv[kMaxMem];
    int status;
};

__device__ __forceinline__ int solve_cta(const SolveArgs& A, const double* sums, double* __restrict__ SY,
                                         double* __restrict__ YY, double* __restrict__ SS, SolveShared& sh,
                                         double* coef_s, bool write_gram, bool comm_ok, int nthreads)
{
    const int m = A.msize, used = A.used, c = A.pend;
    auto ph = [&](int i) { int s = A.oldest + i; return s >= m ? s - m : s; };      // logical (oldest..newest) -> physical slot
    // stage everything the serial solve touches in shared memory, in logical order
    for (int t = threadIdx.x; t < used * used; t += nthreads) {
        const int i = t / used, j = t % used;
        const int pi = ph(i), pj = ph(j);
        double r = __ldcg(SY + pi * m + pj), yy = __ldcg(YY + pi * m + pj);      // (L2: another CTA of a persistent grid may have written them)
        if (c >= 0) {
            if (pj == c) { r = sums[2 * m + pi]; yy = sums[3 * m + pi]; }
            else if (pi == c) yy = sums[3 * m + pj];
        }
        sh.Rm[i][j] = r;
        sh.Yl[i][j] = yy;
    }
    for (int i = threadIdx.x; i < used; i += nthreads) {
        const int pi = ph(i);
        sh.pv[i] = sums[pi];
        sh.qv[i] = sums[m + pi];
        sh.ssv[i] = (pi == c) ? sums[4 * m + 1] : __ldcg(SS + pi);
    }
    for (int j = threadIdx.x; j < 2 * m; j += nthreads) coef_s[j] = 0.0;
    if (write_gram && c >= 0) {              // fold the newest pair's Gram column into the global state
        for (int j = threadIdx.x; j < used; j += nthreads) {
            SY[j * m + c] = sums[2 * m + j];
            const double yy = sums[3 * m + j];
            YY[j * m + c] = yy;
            YY[c * m + j] = yy;
        }
        if (threadIdx.x == 0) SS[c] = sums[4 * m + 1];
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        // warp-parallel solve: lane i owns row i (used <= kMaxMem = 32).  Column-oriented substitutions, one shuffle
        // broadcast + one FMA per lane and step, reciprocal diagonal computed up front (a zero or non-finite pivot
        // yields Inf / NaN exactly as the division would).  ~1 us for m = 10 instead of ~5 us on a single thread.
        const unsigned full = 0xffffffffu;
        const int i = threadIdx.x;
        const bool live = i < used;
        const double gg = sums[4 * m];
        double gamma = 1.0, U = sqrt(gg);
        bool ok = finite_d(gg);
        if (used > 0) {
            gamma = (A.h0 > 0) ? A.h0 : sh.Rm[used - 1][used - 1] / sh.Yl[used - 1][used - 1];
            const double rii = live ? sh.Rm[i][i] : 1.0;
            const double inv = 1.0 / rii;
            double t = live ? sh.pv[i] : 0.0, ui = 0.0;
            for (int j = used - 1; j >= 0; --j) {          // u = R^-1 p   (back substitution)
                const double uj = __shfl_sync(full, t * inv, j);
                if (i == j) ui = uj;
                if (i < j) t = fma(-sh.Rm[i][j], uj, t);
            }
            double acc = 0.0;                               // w = (D + gamma*YY) u - gamma*q0
            for (int j = 0; j < used; ++j) {
                const double uj = __shfl_sync(full, ui, j);
                if (live) acc = fma(sh.Yl[i][j], uj, acc);
            }
            t = live ? rii * ui + gamma * acc - gamma * sh.qv[i] : 0.0;
            double ai = 0.0;
            for (int j = 0; j < used; ++j) {               // a = R^-T w   (forward substitution)
                const double aj = __shfl_sync(full, t * inv, j);
                if (i == j) ai = aj;
                if (live && i > j) t = fma(-sh.Rm[j][i], aj, t);
            }
            const double gb = -gamma * ui;
            double term = 0.0;
            bool fin = true;
            if (live) {
                const int sl = ph(i);
                coef_s[sl] = ai;
                coef_s[m + sl] = gb;
                term = fabs(ai) * sqrt(sh.ssv[i]) + fabs(gb) * sqrt(sh.Yl[i][i]);
                fin = finite_d(ai) && finite_d(gb);
            }
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) term += __shfl_xor_sync(full, term, o);
            U = fabs(gamma) * sqrt(gg) + term;
            ok = ok && __all_sync(full, fin) && finite_d(gamma);
        }
        ok = ok && finite_d(U);
        if (i == 0) {
            coef_s[2 * m] = gamma;
            coef_s[2 * m + 1] = U;
            coef_s[2 * m + 2] = gg;
            int st = ST_ACCEPT;
            if (A.check_nan) {
                if (!ok) st = ST_REJECT_NONFINITE;
                else if (used == 0) { if (U > A.limit) st = ST_REJECT_NONFINITE; }      // d = g: U is the exact norm (stochqn.c:829)
                else if (!(U <= 0.99 * A.limit)) st = ST_NEED_EXACT_NORM;
            }
            if (!comm_ok) st = ST_COMM_TIMEOUT;
            sh.status = st;
        }
    }
    __syncthreads();
    return sh.status;
}

// status word + the three diagnostics to the device word and the mapped host block, then the sequence number
__device__ __forceinline__ void publish_status(int st, const double* coef_s, int m, int* status_dev, volatile int* status_host,
                                               volatile double* info_host, volatile unsigned long long* seq_host,
                                               unsigned long long seq)
{
    *status_dev = st;
    *status_host = st;
    info_host[0] = coef_s[2 * m + 1];
    info_host[1] = coef_s[2 * m];
    info_host[2] = coef_s[2 * m + 2];
    publish_seq(seq_host, seq);
}

__global__ void __launch_bounds__(kThreads)
k2_solve(SolveArgs A, PeerArgs pa, const double* __restrict__ partials, double* sums,
         double* __restrict__ SY, double* __restrict__ YY, double* __restrict__ SS,
         double* __restrict__ coef, int* __restrict__ status_dev, volatile int* status_host,
         volatile double* info_host, volatile unsigned long long* seq_host)
{
    const int m = A.msize;
    const int P = 4 * m + 2;
    if (A.nblocks > 0) { reduce_partials(partials, A.nblocks, P, sums); __syncthreads(); }
    bool comm_ok = true;
    if (pa.world > 1) comm_ok = p2p_allreduce_cta(pa, sums, P);      // sharded: sum the records of all ranks (p2p.cuh)
    if (!A.do_solve) return;
    __shared__ SolveShared sh;
    __shared__ double coef_s[2 * kMaxMem + 3];
    const int st = solve_cta(A, sums, SY, YY, SS, sh, coef_s, true, comm_ok, kThreads);
    for (int j = threadIdx.x; j < 2 * m + 3; j += kThreads) coef[j] = coef_s[j];
    __syncthreads();
    if (threadIdx.x == 0) publish_status(st, coef_s, m, status_dev, status_host, info_host, seq_host, A.seq);
}

// =========================================================================================
// K3: fused combine + update pass.
//   d = gamma*g + sum_j a_j s_j + sum_j (gamma b_j) y_j            (T arithmetic, FMA chains)
//   MODE_OLBFGS : x -= step*d ; s_slot <- -step*d ; grad <- -step*d  (stochqn.c:838,1006-1007)
//   MODE_AVG    : x -= step*d ; x_sum += x ; grad <- d               (stochqn.c:838,1067 / 1191)
//   MODE_DIRONLY: grad <- d, and sum d^2 / non-finite count into partials (exact-norm route)
// Does nothing unless *status_dev == ST_ACCEPT (or `force`), so a rejected direction never
// touches x - the reference's check-before-update order (stochqn.c:825-838).
//
// Work split (picked with tools/kbench_k3.cu): 128-thread CTAs = 2 row-groups x 64 chunk-lanes, 4 CTAs per SM.
// Group 0 forms gamma*g + sum_j a_j s_j, group 1 forms sum_j (gamma b_j) y_j for the same 64 chunks; every load
// of an iteration (up to m rows + g + x [+ x_sum] per thread) is issued before the first use.  The two partials
// meet in shared memory (double-buffered, one barrier per iteration) and each group then performs its output
// job: group 0 updates x (and x_sum), group 1 stores the new s and grad.
// The barrier also orders the read of the slot that is about to be overwritten (it is still a
// valid, oldest pair) and of `grad` before the stores that replace them.
// =========================================================================================
enum : int { MODE_OLBFGS = 0, MODE_AVG = 1, MODE_DIRONLY = 2 };

constexpr int k3Groups = 2;                    // group 0 streams the S rows (+ g, x), group 1 the Y rows
constexpr int k3Lanes = 64;
constexpr int k3Threads = k3Groups * k3Lanes;  // 128-thread CTAs, 4 resident per SM (tools/kbench_k3.cu)
constexpr int k3_min_blocks(int rpg) { return rpg <= 10 ? 4 : rpg <= 16 ? 3 : 2; }

template <typename T, int RPG, int MODE, int VEC>
__global__ void __launch_bounds__(k3Threads, k3_min_blocks(RPG))
k3_combine(const T* g_in, T* grad_out, const T* S_ro, const T* __restrict__ Y,
           T* S_rw, size_t ld, int msize, int used, int new_slot, long long n,
           T* __restrict__ x, T* __restrict__ x_sum, T step, const double* __restrict__ coef,
           const int* __restrict__ status_dev, int force, double* __restrict__ partials)
{
    if (!force && *status_dev != ST_ACCEPT) return;
    const int group = threadIdx.x / k3Lanes, lane = threadIdx.x % k3Lanes;
    const T* rows[RPG];
    T cf[RPG];
    #pragma unroll
    for (int r = 0; r < RPG; ++r) {
        const bool live = r < used;
        const int j = live ? r : (used > 0 ? used - 1 : 0);
        rows[r] = (group == 0 ? S_ro : Y) + (size_t) j * ld;
        cf[r] = live ? (T) coef[group == 0 ? j : msize + j] : (T) 0;
    }
    const T gamma = (T) coef[2 * msize];
    const T nstep = -step;
    double a_dd = 0, a_bad = 0;
    __shared__ __align__(16) T xchg[2][k3Groups][k3Lanes * (VEC > 1 ? VEC : 1)];
    int buf = 0;

    auto one = [&](size_t off, bool valid, auto vtag) {
        constexpr int V = decltype(vtag)::value;
        Pack<T, V> part, xv, xs;
        #pragma unroll
        for (int e = 0; e < V; ++e) part.set(e, (T) 0);
        if (valid) {
            Pack<T, V> rv[RPG], gv;
            // every load of the iteration is issued before the first use; the rows are read through the coherent
            // path because the slot that receives the new s is one of them
            // (dead rows of a partly filled memory are neither loaded nor combined: 0 * Inf would poison d)
            #pragma unroll
            for (int r = 0; r < RPG; ++r) if (r < used) rv[r] = ld_rw<T, V>(rows[r] + off);
            if (group == 0) {
                gv = ld_rw<T, V>(g_in + off);
                if constexpr (MODE != MODE_DIRONLY) xv = ld_rw<T, V>(x + off);
                if constexpr (MODE == MODE_AVG) xs = ld_rw<T, V>(x_sum + off);
                #pragma unroll
                for (int e = 0; e < V; ++e) part.set(e, gamma * gv.get(e));
            }
            #pragma unroll
            for (int r = 0; r < RPG; ++r) {
                if (r < used) {
                    #pragma unroll
                    for (int e = 0; e < V; ++e) part.set(e, fma(cf[r], rv[r].get(e), part.get(e)));
                }
            }
        }
        T* slot = &xchg[buf][group][lane * V];
        #pragma unroll
        for (int e = 0; e < V; ++e) slot[e] = part.get(e);
        __syncthreads();
        if (valid) {
            Pack<T, V> d;
            #pragma unroll
            for (int e = 0; e < V; ++e) d.set(e, xchg[buf][0][lane * V + e] + xchg[buf][1][lane * V + e]);
            if constexpr (MODE == MODE_DIRONLY) {
                if (group == 0) {
                    #pragma unroll
                    for (int e = 0; e < V; ++e) {
                        double de = (double) d.get(e);
                        a_dd = fma(de, de, a_dd);
                        if (!isfinite(de)) a_bad += 1.0;
                    }
                } else st_vec<T, V>(grad_out + off, d);
            } else {
                if (group == 0) {
                    #pragma unroll
                    for (int e = 0; e < V; ++e) xv.set(e, fma(nstep, d.get(e), xv.get(e)));
                    st_vec<T, V>(x + off, xv);
                    if constexpr (MODE == MODE_AVG) {
                        #pragma unroll
                        for (int e = 0; e < V; ++e) xs.set(e, xs.get(e) + xv.get(e));
                        st_vec<T, V>(x_sum + off, xs);
                    }
                } else {
                    if constexpr (MODE == MODE_OLBFGS) {
                        #pragma unroll
                        for (int e = 0; e < V; ++e) d.set(e, nstep * d.get(e));
                        st_vec<T, V>(S_rw + (size_t) new_slot * ld + off, d);
                    }
                    if (grad_out) st_vec<T, V>(grad_out + off, d);
                }
            }
        }
        buf ^= 1;
    };

    const long long nchunks = n / VEC;
    const long long stride = (long long) gridDim.x * k3Lanes;
    for (long long base = (long long) blockIdx.x * k3Lanes; base < nchunks; base += stride) {   // block-uniform trip count
        const long long c = base + lane;
        one((size_t) c * VEC, c < nchunks, std::integral_constant<int, VEC>{});
    }
    if (VEC > 1 && blockIdx.x == 0 && nchunks * VEC < n) {
        const long long i = nchunks * VEC + lane;
        one((size_t) i, i < n, std::integral_constant<int, 1>{});
    }
    if constexpr (MODE == MODE_DIRONLY) {
        double* out = partials + (size_t) blockIdx.x * 2;
        block_reduce<2, k3Threads>(2, [&](int p) { return p == 0 ? a_dd : a_bad; }, [&](int p, double v) { out[p] = v; });
    }
}

// After MODE_DIRONLY decided "accept": x -= step*d with d read back from grad, plus the epilogue.
template <typename T, int MODE, int VEC>
__global__ void __launch_bounds__(kThreads)
k3_apply(T* grad, T* S_rw, size_t ld, int new_slot, long long n,
         T* __restrict__ x, T* __restrict__ x_sum, T step)
{
    const T nstep = -step;
    auto one = [&](size_t off, auto vtag) {
        constexpr int V = decltype(vtag)::value;
        Pack<T, V> d = ld_rw<T, V>(grad + off);
        Pack<T, V> xv = ld_rw<T, V>(x + off);
        #pragma unroll
        for (int e = 0; e < V; ++e) xv.set(e, fma(nstep, d.get(e), xv.get(e)));
        st_vec<T, V>(x + off, xv);
        if constexpr (MODE == MODE_OLBFGS) {
            Pack<T, V> sn;
            #pragma unroll
            for (int e = 0; e < V; ++e) sn.set(e, nstep * d.get(e));
            st_vec<T, V>(S_rw + (size_t) new_slot * ld + off, sn);
            st_vec<T, V>(grad + off, sn);
        } else {
            Pack<T, V> xs = ld_rw<T, V>(x_sum + off);
            #pragma unroll
            for (int e = 0; e < V; ++e) xs.set(e, xs.get(e) + xv.get(e));
            st_vec<T, V>(x_sum + off, xs);
        }
    };
    const long long nchunks = n / VEC;
    const long long stride = (long long) gridDim.x * kThreads;
    for (long long c = (long long) blockIdx.x * kThreads + threadIdx.x; c < nchunks; c += stride)
        one((size_t) c * VEC, std::integral_constant<int, VEC>{});
    if (VEC > 1 && blockIdx.x == 0) {
        const long long i = nchunks * VEC + threadIdx.x;
        if (i < n) one((size_t) i, std::integral_constant<int, 1>{});
    }
}

// =========================================================================================
// K4: correction-pair construction with the curvature dots.
//   PAIR_GRAD_DIFF: y_slot = g - g_prev (+ y_reg*s)   (stochqn.c:915-923)
//   PAIR_COPY     : y_slot = src (Hessian-vector product, stochqn.c:964)
//   PAIR_DOTS_ONLY: y_slot already holds y (Fisher product)
// partial record: [0] s'y  [1] s's   (the two dots of stochqn.c:892)
// Optional fused copies on acceptance are done by separate tiny kernels (the decision is
// taken on the host after the dots are known).
// =========================================================================================
enum : int { PAIR_GRAD_DIFF = 0, PAIR_COPY = 1, PAIR_DOTS_ONLY = 2 };

// =========================================================================================
// K4 with the finalisation folded in ("last block done"): every CTA writes its 2-value record and takes a ticket;
// the CTA that draws the last ticket sums the records in CTA order (deterministic) and publishes s'y, s's to the
// mapped host block.  Saves the k_finalize launch whenever the optimizer is not sharded (any n).
// =========================================================================================
__device__ __forceinline__ void last_block_publish2(const double* __restrict__ partials, double* __restrict__ sums,
                                                    unsigned int* ticket, volatile double* host_out,
                                                    volatile unsigned long long* seq_host, unsigned long long seq,
                                                    const PeerArgs& pa)
{
    __shared__ int is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp < 2) {
        double v = 0;
        for (unsigned b = lane; b < gridDim.x; b += 32) v += __ldcg(partials + (size_t) b * 2 + warp);
        v = warp_sum(v);
        if (lane == 0) sums[warp] = v;
    }
    __syncthreads();
    // sharded: the same CTA exchanges the two values with the other ranks over peer memory (no k_finalize launch)
    bool ok = true;
    if (pa.world > 1) ok = p2p_allreduce_cta(pa, sums, 2);
    if (threadIdx.x < 2) host_out[threadIdx.x] = ok ? sums[threadIdx.x] : __longlong_as_double(0x7ff8000000000000ll);
    __syncthreads();
    if (threadIdx.x == 0) { *ticket = 0; publish_seq(seq_host, seq); }
}


template <typename T, int KIND, int VEC>
__global__ void __launch_bounds__(kThreads)
k4_pair(const T* __restrict__ a, const T* __restrict__ b, const T* __restrict__ s, T* __restrict__ y,
        T y_reg, long long n, double* __restrict__ partials,
        unsigned int* ticket, double* __restrict__ sums, volatile double* host_out,
        volatile unsigned long long* seq_host, unsigned long long seq, PeerArgs pa)
{
    double a_sy = 0, a_ss = 0;
    auto one = [&](size_t off, auto vtag) {
        constexpr int V = decltype(vtag)::value;
        Pack<T, V> sv = ld_stream<T, V>(s + off);
        Pack<T, V> yv;
        if constexpr (KIND == PAIR_GRAD_DIFF) {
            Pack<T, V> av = ld_stream<T, V>(a + off), bv = ld_stream<T, V>(b + off);
            #pragma unroll
            for (int e = 0; e < V; ++e) {
                T t = av.get(e) - bv.get(e);
                if (y_reg > (T) 0) t = fma(y_reg, sv.get(e), t);
                yv.set(e, t);
            }
            st_vec<T, V>(y + off, yv);
        } else if constexpr (KIND == PAIR_COPY) {
            yv = ld_stream<T, V>(a + off);
            st_vec<T, V>(y + off, yv);
        } else {
            yv = ld_rw<T, V>(y + off);
        }
        #pragma unroll
        for (int e = 0; e < V; ++e) {
            double se = (double) sv.get(e), ye = (double) yv.get(e);
            a_sy = fma(se, ye, a_sy);
            a_ss = fma(se, se, a_ss);
        }
    };
    const long long nchunks = n / VEC;
    const long long stride = (long long) gridDim.x * kThreads;
    for (long long c = (long long) blockIdx.x * kThreads + threadIdx.x; c < nchunks; c += stride)
        one((size_t) c * VEC, std::integral_constant<int, VEC>{});
    if (VEC > 1 && blockIdx.x == 0) {
        const long long i = nchunks * VEC + threadIdx.x;
        if (i < n) one((size_t) i, std::integral_constant<int, 1>{});
    }
    double* out = partials + (size_t) blockIdx.x * 2;
    block_reduce<2>(2, [&](int p) { return p == 0 ? a_sy : a_ss; }, [&](int p, double v) { out[p] = v; });
    if (ticket) last_block_publish2(partials, sums, ticket, host_out, seq_host, seq, pa);
}

// Sum `count`-wide partial records over CTAs into `sums` (device), across ranks when sharded (p2p.cuh), and, when
// asked, into mapped host memory followed by the sequence number the host polls.  One CTA.
// Used for the 2-value records of K4 / MODE_DIRONLY and the k Fisher dots.
__global__ void __launch_bounds__(kThreads)
k_finalize(const double* __restrict__ partials, int nblocks, int count, double* sums, PeerArgs pa,
           volatile double* host_out, volatile unsigned long long* seq_host, unsigned long long seq)
{
    reduce_partials(partials, nblocks, count, sums);
    __syncthreads();
    bool ok = true;
    if (pa.world > 1) ok = p2p_allreduce_cta(pa, sums, count);
    if (host_out) {
        for (int p = threadIdx.x; p < count; p += kThreads) host_out[p] = ok ? sums[p] : __longlong_as_double(0x7ff8000000000000ll);
        __syncthreads();
        if (threadIdx.x == 0) publish_seq(seq_host, seq);
    }
}

// The same sums for WIDE records on one GPU (the k Fisher dots: 100 values x 296 records): one warp per entry, eight entries per
// CTA, as many CTAs as it takes - a single CTA walking 30 000 strided values was 27 us of the adaQN pair iteration.
__global__ void __launch_bounds__(kThreads)
k_finalize_wide(const double* __restrict__ partials, int nblocks, int count, double* __restrict__ sums)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int p = blockIdx.x * kWarps + warp;
    if (p >= count) return;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    int b = lane;
    for (; b + 96 < nblocks; b += 128) {
        a0 += __ldcg(partials + (size_t) b * count + p);         a1 += __ldcg(partials + (size_t) (b + 32) * count + p);
        a2 += __ldcg(partials + (size_t) (b + 64) * count + p);  a3 += __ldcg(partials + (size_t) (b + 96) * count + p);
    }
    for (; b < nblocks; b += 32) a0 += __ldcg(partials + (size_t) b * count + p);
    const double v = warp_sum((a0 + a1) + (a2 + a3));
    if (lane == 0) sums[p] = v;
}

// device -> mapped-host copy of a few doubles (after a library all-reduce landed them in `sums`)
__global__ void k_publish(const double* __restrict__ sums, int count, volatile double* host_out,
                          volatile unsigned long long* seq_host, unsigned long long seq)
{
    if (threadIdx.x < count) host_out[threadIdx.x] = sums[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) publish_seq(seq_host, seq);
}

// stand-alone small all-reduce (sum, fp64, count <= kBoxCap) over peer memory: callbacks ride on it (halo exchange)
__global__ void __launch_bounds__(kThreads)
k_p2p_allreduce(double* buf, int count, PeerArgs pa, int* __restrict__ error_flag)
{
    // sticky: once an exchange of this communicator has timed out, the later ones do not wait another 20 s each
    if (error_flag && *reinterpret_cast<volatile int*>(error_flag)) return;
    if (!p2p_allreduce_cta(pa, buf, count) && threadIdx.x == 0 && error_flag) *error_flag = 1;
}

// Gram bookkeeping for a slot that was zeroed by a rejected pair (quirk Q1): the slot now holds
// s = y = 0, so every product with it is 0.
__global__ void k_gram_zero_slot(double* SY, double* YY, double* SS, int m, int c)
{
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
        SY[j * m + c] = 0; SY[c * m + j] = 0;
        YY[j * m + c] = 0; YY[c * m + j] = 0;
    }
    if (threadIdx.x == 0) SS[c] = 0;
}

// ---- iterate averaging helpers (SQN / adaQN) ---------------------------------------------
enum : int { AVG_ADD = 0, AVG_ARCHIVE = 1, AVG_S_VECTOR = 2, AVG_SCALE = 3 };
// AVG_ADD      : x_sum += x                                   (stochqn.c:268-284; used when a step was rejected, Q7)
// AVG_ARCHIVE  : x_avg_prev = x_sum*inv ; x_sum = 0           (stochqn.c:1080-1081: average_from_sum + archive_x_avg)
// AVG_S_VECTOR : x_sum *= inv ; s_slot = x_sum - x_avg_prev   (stochqn.c:861-870)
// AVG_SCALE    : x_sum *= inv                                 (stochqn.c:1229)
template <typename T, int OP, int VEC>
__global__ void __launch_bounds__(kThreads)
k_avg(T* __restrict__ x_sum, T* __restrict__ other, T* __restrict__ s_slot, T inv, long long n)
{
    auto one = [&](size_t off, auto vtag) {
        constexpr int V = decltype(vtag)::value;
        Pack<T, V> xs = ld_rw<T, V>(x_sum + off);
        if constexpr (OP == AVG_ADD) {
            Pack<T, V> xv = ld_rw<T, V>(other + off);
            #pragma unroll
            for (int e = 0; e < V; ++e) xs.set(e, xs.get(e) + xv.get(e));
            st_vec<T, V>(x_sum + off, xs);
        } else if constexpr (OP == AVG_ARCHIVE) {
            Pack<T, V> z;
            #pragma unroll
            for (int e = 0; e < V; ++e) { xs.set(e, xs.get(e) * inv); z.set(e, (T) 0); }
            st_vec<T, V>(other + off, xs);
            st_vec<T, V>(x_sum + off, z);
        } else if constexpr (OP == AVG_S_VECTOR) {
            Pack<T, V> xp = ld_rw<T, V>(other + off), sv;
            #pragma unroll
            for (int e = 0; e < V; ++e) { xs.set(e, xs.get(e) * inv); sv.set(e, xs.get(e) - xp.get(e)); }
            st_vec<T, V>(x_sum + off, xs);
            st_vec<T, V>(s_slot + off, sv);
        } else {
            #pragma unroll
            for (int e = 0; e < V; ++e) xs.set(e, xs.get(e) * inv);
            st_vec<T, V>(x_sum + off, xs);
        }
    };
    const long long nchunks = n / VEC;
    const long long stride = (long long) gridDim.x * kThreads;
    for (long long c = (long long) blockIdx.x * kThreads + threadIdx.x; c < nchunks; c += stride)
        one((size_t) c * VEC, std::integral_constant<int, VEC>{});
    if (VEC > 1 && blockIdx.x == 0) {
        const long long i = nchunks * VEC + threadIdx.x;
        if (i < n) one((size_t) i, std::integral_constant<int, 1>{});
    }
}


// =========================================================================================
// Empirical-Fisher product (stochqn.c:936-952):  y = F' (F s) / k  over the first k ring rows.
//   KF1: t_r = F_r's for 4*RPG rows per launch (record [k], entry r) - the work split of K1: 4 row-groups x 64
//        chunk-lanes, two chunks in flight per thread, rows loaded with L1::no_allocate, s shared through L1
//   KF2: y = (1/k) sum_r t_r F_r in ONE launch - the work split of K3: 128-thread CTAs = 2 row-groups x 64 lanes, each
//        group streams half of the rows (RB loads in flight), the halves meet in shared memory, y is written once
// Two sweeps of F (2k vector transfers + one read of s per KF1 launch + one write of y): HBM-bound skinny GEMV pair.
// =========================================================================================
template <typename T, int RPG, int VEC>
__global__ void __launch_bounds__(kThreads, 2)
kf1_rowdots(const T* __restrict__ F, size_t ld, int r0, int rows, const T* __restrict__ s, long long n,
            int rec, double* __restrict__ partials)
{
    const int group = threadIdx.x / kLanes, lane = threadIdx.x % kLanes;
    const T* rp[RPG];
    #pragma unroll
    for (int r = 0; r < RPG; ++r) {
        int v = group * RPG + r;
        if (v >= rows) v = rows - 1;                 // clamped duplicate: dropped at the end
        rp[r] = F + (size_t) (r0 + v) * ld;
    }
    double acc[RPG];
    #pragma unroll
    for (int r = 0; r < RPG; ++r) acc[r] = 0;
    auto many = [&](const long long (&c)[kUnroll], auto vtag) {
        constexpr int V = decltype(vtag)::value;
        Pack<T, V> sv[kUnroll], rv[kUnroll][RPG];
        #pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (c[u] >= 0) {
                const size_t off = (size_t) c[u] * V;
                sv[u] = ld_stream<T, V>(s + off);
                #pragma unroll
                for (int r = 0; r < RPG; ++r) rv[u][r] = ld_row<T, V>(rp[r] + off);
            }
        }
        #pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (c[u] >= 0) {
                #pragma unroll
                for (int r = 0; r < RPG; ++r) {
                    #pragma unroll
                    for (int e = 0; e < V; ++e) acc[r] = fma((double) rv[u][r].get(e), (double) sv[u].get(e), acc[r]);
                }
            }
        }
    };
    const long long nchunks = n / VEC;
    const long long stride = (long long) gridDim.x * kLanes;
    for (long long c0 = (long long) blockIdx.x * kLanes + lane; c0 < nchunks; c0 += stride * kUnroll) {
        long long c[kUnroll];
        #pragma unroll
        for (int u = 0; u < kUnroll; ++u) { c[u] = c0 + u * stride; if (c[u] >= nchunks) c[u] = -1; }
        many(c, std::integral_constant<int, VEC>{});
    }
    if (VEC > 1 && blockIdx.x == 0) {               // scalar tail: n not a multiple of VEC
        const long long i = nchunks * VEC + lane;
        long long c[kUnroll];
        #pragma unroll
        for (int u = 0; u < kUnroll; ++u) c[u] = -1;
        if (i < n) { c[0] = i; many(c, std::integral_constant<int, 1>{}); }
    }
    __shared__ double red[kWarps][RPG];
    const int warp = threadIdx.x >> 5, wl = threadIdx.x & 31;
    #pragma unroll
    for (int r = 0; r < RPG; ++r) {
        const double v = warp_sum(acc[r]);
        if (wl == 0) red[warp][r] = v;
    }
    __syncthreads();
    double* out = partials + (size_t) blockIdx.x * rec + r0;
    constexpr int WPG = kWarps / kGroups;
    for (int t = threadIdx.x; t < kGroups * RPG; t += kThreads) {
        const int gi = t / RPG, r = t % RPG, vrow = gi * RPG + r;
        if (vrow >= rows) continue;
        double v = 0;
        #pragma unroll
        for (int w = 0; w < WPG; ++w) v += red[gi * WPG + w][r];
        out[vrow] = v;
    }
}

template <typename T, int RB, int VEC>
__global__ void __launch_bounds__(k3Threads, 4)
kf2_combine(const T* __restrict__ F, size_t ld, int k, const double* __restrict__ t, double inv_k,
            T* __restrict__ y, long long n)
{
    extern __shared__ __align__(16) unsigned char kf2_smem[];
    T* cs = reinterpret_cast<T*>(kf2_smem);                        // buffer_y in storage precision (stochqn.c:946-947)
    for (int i = threadIdx.x; i < k; i += k3Threads) cs[i] = (T) t[i];
    __shared__ __align__(16) T xchg[2][k3Lanes * (VEC > 1 ? VEC : 1)];
    __syncthreads();
    const int group = threadIdx.x / k3Lanes, lane = threadIdx.x % k3Lanes;
    const int half = (k + 1) / 2;
    const int r_begin = group == 0 ? 0 : half, r_end = group == 0 ? half : k;
    const T alpha = (T) inv_k;
    int buf = 0;
    auto one = [&](size_t off, bool valid, auto vtag) {
        constexpr int V = decltype(vtag)::value;
        Pack<T, V> acc;
        #pragma unroll
        for (int e = 0; e < V; ++e) acc.set(e, (T) 0);
        if (valid) {
            for (int rb = r_begin; rb < r_end; rb += RB) {
                Pack<T, V> fv[RB];
                #pragma unroll
                for (int q = 0; q < RB; ++q) {
                    const int r = rb + q < r_end ? rb + q : r_end - 1;
                    fv[q] = ld_row<T, V>(F + (size_t) r * ld + off);
                }
                #pragma unroll
                for (int q = 0; q < RB; ++q) {
                    const T cq = rb + q < r_end ? cs[rb + q] : (T) 0;
                    #pragma unroll
                    for (int e = 0; e < V; ++e) acc.set(e, fma(cq, fv[q].get(e), acc.get(e)));
                }
            }
        }
        if (group == 1) {
            #pragma unroll
            for (int e = 0; e < V; ++e) xchg[buf][lane * V + e] = acc.get(e);
        }
        __syncthreads();
        if (group == 0 && valid) {
            Pack<T, V> yv;
            #pragma unroll
            for (int e = 0; e < V; ++e) yv.set(e, alpha * (acc.get(e) + xchg[buf][lane * V + e]));
            st_vec<T, V>(y + off, yv);
        }
        buf ^= 1;
    };
    const long long nchunks = n / VEC;
    const long long stride = (long long) gridDim.x * k3Lanes;
    for (long long base = (long long) blockIdx.x * k3Lanes; base < nchunks; base += stride) {     // block-uniform trip count
        const long long c = base + lane;
        one((size_t) c * VEC, c < nchunks, std::integral_constant<int, VEC>{});
    }
    if (VEC > 1 && blockIdx.x == 0 && nchunks * VEC < n) {
        const long long i = nchunks * VEC + lane;
        one((size_t) i, i < n, std::integral_constant<int, 1>{});
    }
}

}  // namespace sqn

#include "kernels_small.cuh"
#include "kernels_adaqn.cuh"
#include "kernels_loop.cuh"
#include "kernels_fit.cuh"
