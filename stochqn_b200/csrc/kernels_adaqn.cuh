// kernels_adaqn.cuh - the adaQN flavour of the step (included at the end of kernels.cuh).
//
// The reference's take_step for adaQN (stochqn.c:808-822 with 720-783) first updates the AdaGrad / RMSProp
// accumulator G, then
//   no pairs : d = g / sqrt(G + eps)
//   pairs    : "H0" = h = g / sqrt(G + eps) (the RESCALED GRADIENT, quirk Q2) and the two-loop multiplies by it
//              elementwise, i.e. H0 = diag(h).
// Compact form with a diagonal H0 = diag(h):   p = S'g,  R = upper(S'Y),  D = diag(S'Y),
//     u = R^-1 p ;  b = -u ;  a = R^-T [ D u + Y' diag(h) (Y u - g) ] ;   d = h.(g + Y b) + S a
// h depends on the current gradient, so nothing that involves it can be kept between steps.  Instead of forming
// the weighted Gram matrix W = Y' diag(h) Y (m(m+1)/2 weighted dots per step), the product (W u - Y'(h.g)) is
// taken as  Y' [ h . (Y u - g) ]:  m dots over a vector that every thread rebuilds from the m values of Y it has
// just loaded.  Per step (pairs present) the sweeps are
//   KA1  ka1_dots     g, G, S (+ y_c of the pending pair):  p, pending Gram column, sum h^2; writes G (and the
//                     Fisher ring row): m+3 reads, 1-2 writes
//   KAu  ka_solve_u   one CTA: sums, (exchange,) Gram fold, u = R^-1 p  -> coef[m..2m) = b
//   KA2  ka2_wdots    g, G, Y: w_j = y_j' [h . (Y u - g)], sum (h.(Yu-g))^2 : m+2 reads
//   KAa  ka_solve_a   one CTA: sums, (exchange,) a = R^-T (D u + w), bound on ||d||, accept flag
//   KA3  ka3_combine  g, G, S, Y, x, x_sum: d, x -= step*d, x_sum += x, grad <- d: 2m+4 reads, 3 writes
// = 4m+9 reads + 5 writes (54 n-vectors for m = 10) against ~140 for the reference's formulation.
#pragma once

namespace sqn {

template <typename T>
__device__ __forceinline__ T ada_accumulate(T g, T G, T w_old, T w_new, bool rms)
{
    // stochqn.c:738 / 745
    return rms ? (w_old * G + w_new * (g * g)) : (G + g * g);
}

// record of KA1 (m = mem_size):  [0,m) p_j = s_j'g   [m,2m) s_j'y_c (pending pair c)
//                                [2m] sum h^2   [2m+1] s_c's_c   [2m+2] y_c'y_c   [2m+3] sum (h g)^2
// record of KA2:                 [0,m) w_j   [m] sum (h.(Yu-g))^2
__host__ __device__ constexpr int ka1_record(int m) { return 2 * m + 4; }
__host__ __device__ constexpr int ka2_record(int m) { return m + 1; }

// KA1: same work split as K1 (4 row-groups x 64 chunk-lanes, two chunks in flight per thread), rows = S rows only.
// The G update, the Fisher ring write (stochqn.c:581-587: the RAW gradient) and the scalars cost an fp64 divide +
// square root per element: that duty ROTATES over the four groups from one loop iteration to the next (a warp-uniform
// choice), so no warp carries more than a quarter of it (with a fixed lead group its two warps were the critical
// path: 4.9 TB/s).  Every group reads g (and y_c) for its dots - one DRAM read, then L1/L2 hits.
template <typename T, int RPG, bool PENDING, int VEC>
__global__ void __launch_bounds__(kThreads, 2)
ka1_dots(const T* __restrict__ g, T* __restrict__ G, const T* __restrict__ S, size_t ld, int msize, int used,
         const T* __restrict__ sc_row, const T* __restrict__ yc_row, long long n, T* __restrict__ fisher_row,
         T scal_reg, T rmsprop_weight, double* __restrict__ partials)
{
    const int group = threadIdx.x / kLanes, lane = threadIdx.x % kLanes;
    const bool rms = (rmsprop_weight > (T) 0 && rmsprop_weight < (T) 1);
    const T w_new = (T) 1 - rmsprop_weight;
    const T* rows[RPG];
    #pragma unroll
    for (int r = 0; r < RPG; ++r) {
        int v = group * RPG + r;
        if (v >= used) v = used > 0 ? used - 1 : 0;
        rows[r] = S + (size_t) v * ld;
    }
    double a_p[RPG], a_c[RPG];
    double a_hh = 0, a_hg = 0, a_ss = 0, a_yy = 0;
    #pragma unroll
    for (int r = 0; r < RPG; ++r) { a_p[r] = 0; a_c[r] = 0; }

    auto many = [&](const long long (&c)[kUnroll], bool lead, auto vtag) {
        constexpr int V = decltype(vtag)::value;
        Pack<T, V> gv[kUnroll], Gv[kUnroll], yc[kUnroll], sc[kUnroll], rv[kUnroll][RPG];
        #pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (c[u] >= 0) {
                const size_t off = (size_t) c[u] * V;
                gv[u] = ld_stream<T, V>(g + off);
                if (lead) Gv[u] = ld_rw<T, V>(G + off);
                if constexpr (PENDING) {
                    yc[u] = ld_stream<T, V>(yc_row + off);
                    if (lead) sc[u] = ld_stream<T, V>(sc_row + off);
                }
                if (used > 0) {
                    #pragma unroll
                    for (int r = 0; r < RPG; ++r) rv[u][r] = ld_row<T, V>(rows[r] + off);
                }
            }
        }
        #pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (c[u] >= 0) {
                const size_t off = (size_t) c[u] * V;
                if (lead) {
                    #pragma unroll
                    for (int e = 0; e < V; ++e) Gv[u].set(e, ada_accumulate<T>(gv[u].get(e), Gv[u].get(e), rmsprop_weight, w_new, rms));
                    st_vec<T, V>(G + off, Gv[u]);
                    if (fisher_row) st_vec<T, V>(fisher_row + off, gv[u]);
                    #pragma unroll
                    for (int e = 0; e < V; ++e) {
                        const T h = gv[u].get(e) / sqrt(Gv[u].get(e) + scal_reg);       // stochqn.c:778 / 781
                        const double he = (double) h, ge = (double) gv[u].get(e);
                        a_hh = fma(he, he, a_hh);
                        const double t = he * ge;
                        a_hg = fma(t, t, a_hg);
                        if constexpr (PENDING) {
                            const double se = (double) sc[u].get(e), ye = (double) yc[u].get(e);
                            a_ss = fma(se, se, a_ss);
                            a_yy = fma(ye, ye, a_yy);
                        }
                    }
                }
                if (used > 0) {
                    #pragma unroll
                    for (int r = 0; r < RPG; ++r) {
                        #pragma unroll
                        for (int e = 0; e < V; ++e) {
                            const double re = (double) rv[u][r].get(e);
                            a_p[r] = fma(re, (double) gv[u].get(e), a_p[r]);
                            if constexpr (PENDING) a_c[r] = fma(re, (double) yc[u].get(e), a_c[r]);
                        }
                    }
                }
            }
        }
    };
    const long long nchunks = n / VEC;
    const long long stride = (long long) gridDim.x * kLanes;
    int turn = 0;
    for (long long c0 = (long long) blockIdx.x * kLanes + lane; c0 < nchunks; c0 += stride * kUnroll) {
        long long c[kUnroll];
        #pragma unroll
        for (int u = 0; u < kUnroll; ++u) { c[u] = c0 + u * stride; if (c[u] >= nchunks) c[u] = -1; }
        many(c, group == turn, std::integral_constant<int, VEC>{});
        turn = (turn + 1) & (kGroups - 1);
    }
    if (VEC > 1 && blockIdx.x == 0) {               // scalar tail
        const long long i = nchunks * VEC + lane;
        long long c[kUnroll];
        #pragma unroll
        for (int u = 0; u < kUnroll; ++u) c[u] = -1;
        if (i < n) { c[0] = i; many(c, group == 0, std::integral_constant<int, 1>{}); }
    }

    constexpr int NA = 2 * RPG + 4;
    __shared__ double red[kWarps][NA];
    const int warp = threadIdx.x >> 5, wl = threadIdx.x & 31;
    #pragma unroll
    for (int p = 0; p < NA; ++p) {
        double v = p < RPG ? a_p[p < RPG ? p : 0] : p < 2 * RPG ? a_c[(p - RPG) < RPG ? (p - RPG) : 0]
                 : p == 2 * RPG ? a_hh : p == 2 * RPG + 1 ? a_ss : p == 2 * RPG + 2 ? a_yy : a_hg;
        v = warp_sum(v);
        if (wl == 0) red[warp][p] = v;
    }
    __syncthreads();
    const int P = ka1_record(msize);
    double* out = partials + (size_t) blockIdx.x * P;
    constexpr int WPG = kWarps / kGroups;
    for (int t = threadIdx.x; t < kGroups * NA; t += kThreads) {
        const int gi = t / NA, p = t % NA;
        double v = 0;
        #pragma unroll
        for (int w = 0; w < WPG; ++w) v += red[gi * WPG + w][p];
        if (p >= 2 * RPG) {                          // scalars: every group holds a share (rotating duty)
            if (gi == 0) {
                #pragma unroll
                for (int q = 1; q < kGroups; ++q) {
                    #pragma unroll
                    for (int w = 0; w < WPG; ++w) v += red[q * WPG + w][p];
                }
                out[2 * msize + (p - 2 * RPG)] = v;
            }
            continue;
        }
        const int r = p % RPG, j = gi * RPG + r;
        if (j >= used) continue;
        if (p < RPG) out[j] = v;
        else if (PENDING) out[msize + j] = v;
    }
    for (int t = threadIdx.x; t < 2 * msize; t += kThreads) {       // entries this launch owns but did not compute
        const int k = t / msize, j = t % msize;
        if (j >= used || (!PENDING && k == 1)) out[t] = 0.0;
    }
}

// Upper-triangular solve shared by the two small adaQN kernels: u = R^-1 p in logical order (oldest..newest).
struct AdaStage {
    double Rm[kMaxMem][kMaxMem + 1];
    double pv[kMaxMem], u[kMaxMem], w[kMaxMem], av[kMaxMem], ssv[kMaxMem], yyv[kMaxMem];
};

__device__ __forceinline__ void ada_stage_and_solve_u(const SolveArgs& A, const double* sums1, const double* SY, const double* SS,
                                                      const double* YY, AdaStage& st)
{
    const int m = A.msize, used = A.used;
    auto ph = [&](int i) { int s = A.oldest + i; return s >= m ? s - m : s; };
    for (int t = threadIdx.x; t < used * used; t += kThreads) {
        const int i = t / used, j = t % used;
        st.Rm[i][j] = SY[ph(i) * m + ph(j)];
    }
    for (int i = threadIdx.x; i < used; i += kThreads) {
        st.pv[i] = sums1[ph(i)];
        st.ssv[i] = SS[ph(i)];
        st.yyv[i] = YY[ph(i) * m + ph(i)];
    }
    __syncthreads();
    if (threadIdx.x < 32) {                 // warp-parallel back substitution: lane i owns row i (see solve_cta, kernels.cuh)
        const int i = threadIdx.x;
        const bool live = i < used;
        const double inv = 1.0 / (live ? st.Rm[i][i] : 1.0);
        double t = live ? st.pv[i] : 0.0;
        for (int j = used - 1; j >= 0; --j) {
            const double uj = __shfl_sync(0xffffffffu, t * inv, j);
            if (i == j) st.u[i] = uj;
            if (i < j) t = fma(-st.Rm[i][j], uj, t);
        }
    }
    __syncthreads();
}

// KAu: sums of KA1 -> Gram fold -> u.  Writes coef[m + slot] = b = -u (physical slots) for KA2 / KA3.
// With no pairs it also takes the accept decision (d = h, ||d|| = sqrt(sum h^2) exactly; stochqn.c:808-812, 825-835).
__global__ void __launch_bounds__(kThreads)
ka_solve_u(SolveArgs A, PeerArgs pa, const double* __restrict__ partials, double* sums,
           double* __restrict__ SY, double* __restrict__ YY, double* __restrict__ SS,
           double* __restrict__ coef, int* __restrict__ status_dev, volatile int* status_host,
           volatile double* info_host, volatile unsigned long long* seq_host)
{
    const int m = A.msize, used = A.used;
    const int P = ka1_record(m);
    if (A.nblocks > 0) { reduce_partials(partials, A.nblocks, P, sums); __syncthreads(); }
    bool comm_ok = true;
    if (pa.world > 1) comm_ok = p2p_allreduce_cta(pa, sums, P);
    if (!A.do_solve) return;
    if (A.pend >= 0) {
        const int c = A.pend;
        for (int j = threadIdx.x; j < used; j += kThreads) SY[j * m + c] = sums[m + j];
        if (threadIdx.x == 0) { SS[c] = sums[2 * m + 1]; YY[c * m + c] = sums[2 * m + 2]; }
        __syncthreads();
    }
    __shared__ AdaStage st;
    ada_stage_and_solve_u(A, sums, SY, SS, YY, st);
    auto ph = [&](int i) { int s = A.oldest + i; return s >= m ? s - m : s; };
    for (int j = threadIdx.x; j < 2 * m; j += kThreads) coef[j] = 0.0;
    __syncthreads();
    for (int i = threadIdx.x; i < used; i += kThreads) coef[m + ph(i)] = -st.u[i];
    if (threadIdx.x != 0) return;
    if (used == 0) {
        const double hh = sums[2 * m];
        const double U = sqrt(hh);
        int stt = ST_ACCEPT;
        if (A.check_nan && (!finite_d(hh) || !(U <= A.limit))) stt = ST_REJECT_NONFINITE;
        if (!comm_ok) stt = ST_COMM_TIMEOUT;
        coef[2 * m] = 1.0; coef[2 * m + 1] = U; coef[2 * m + 2] = hh;
        *status_dev = stt;
        *status_host = stt;
        info_host[0] = U; info_host[1] = 1.0; info_host[2] = hh;
        publish_seq(seq_host, A.seq);
    } else if (!comm_ok) {
        *status_dev = ST_COMM_TIMEOUT;          // KAa reports it
    } else {
        *status_dev = ST_ACCEPT;
    }
}

// KA2: w_j = y_j' [ h . (Y u - g) ] for every pair, and sum (h.(Yu-g))^2.  Every thread streams ALL the Y rows of
// its chunk (it needs them all to rebuild Y u), 128-thread CTAs.  coef[m + j] holds b_j = -u_j.
constexpr int kaThreads = 128;
constexpr int ka_min_blocks(int mmax) { return mmax <= 12 ? 3 : mmax <= 16 ? 2 : 1; }

template <typename T, int MMAX, int VEC>
__global__ void __launch_bounds__(kaThreads, ka_min_blocks(MMAX))
ka2_wdots(const T* __restrict__ g, const T* __restrict__ G, const T* __restrict__ Y, size_t ld, int msize, int used,
          long long n, T scal_reg, const double* __restrict__ coef, double* __restrict__ partials)
{
    T cu[MMAX];
    #pragma unroll
    for (int j = 0; j < MMAX; ++j) cu[j] = (j < used) ? (T) (-coef[msize + j]) : (T) 0;       // u_j
    double acc[MMAX], a_tt = 0;
    #pragma unroll
    for (int j = 0; j < MMAX; ++j) acc[j] = 0;

    auto one = [&](size_t off, auto vtag) {
        constexpr int V = decltype(vtag)::value;
        Pack<T, V> gv = ld_stream<T, V>(g + off), Gv = ld_stream<T, V>(G + off), yv[MMAX], t;
        #pragma unroll
        for (int j = 0; j < MMAX; ++j) if (j < used) yv[j] = ld_row<T, V>(Y + (size_t) j * ld + off);
        #pragma unroll
        for (int e = 0; e < V; ++e) t.set(e, -gv.get(e));
        #pragma unroll
        for (int j = 0; j < MMAX; ++j) {
            if (j < used) {
                #pragma unroll
                for (int e = 0; e < V; ++e) t.set(e, fma(cu[j], yv[j].get(e), t.get(e)));
            }
        }
        double ht[V];
        #pragma unroll
        for (int e = 0; e < V; ++e) {
            const T h = gv.get(e) / sqrt(Gv.get(e) + scal_reg);
            ht[e] = (double) (h * t.get(e));
            a_tt = fma(ht[e], ht[e], a_tt);
        }
        #pragma unroll
        for (int j = 0; j < MMAX; ++j) {
            if (j < used) {
                #pragma unroll
                for (int e = 0; e < V; ++e) acc[j] = fma((double) yv[j].get(e), ht[e], acc[j]);
            }
        }
    };
    const long long nchunks = n / VEC;
    const long long stride = (long long) gridDim.x * kaThreads;
    for (long long c = (long long) blockIdx.x * kaThreads + threadIdx.x; c < nchunks; c += stride)
        one((size_t) c * VEC, std::integral_constant<int, VEC>{});
    if (VEC > 1 && blockIdx.x == 0) {
        const long long i = nchunks * VEC + threadIdx.x;
        if (i < n) one((size_t) i, std::integral_constant<int, 1>{});
    }
    const int P = ka2_record(msize);
    double* out = partials + (size_t) blockIdx.x * P;
    block_reduce<MMAX + 1, kaThreads>(MMAX + 1,
        [&](int p) -> double { return p < MMAX ? acc[p < MMAX ? p : 0] : a_tt; },
        [&](int p, double v) {
            if (p == MMAX) out[msize] = v;
            else if (p < msize) out[p] = (p < used) ? v : 0.0;
        });
}

// KAa: sums of KA2 -> a = R^-T (D u + w), bound on ||d||, accept flag.  `sums1` = the (already reduced) KA1 sums.
// coef layout: [0,m) a, [m,2m) b (left as KAu wrote them), [2m] 1, [2m+1] U, [2m+2] sum (h.(Yu-g))^2.
//   ||d|| = ||-h.(Yu-g) + S a|| <= sqrt(sum (h.(Yu-g))^2) + sum_i |a_i| ||s_i||   (no cancellation)
__global__ void __launch_bounds__(kThreads)
ka_solve_a(SolveArgs A, PeerArgs pa, const double* __restrict__ partials, const double* __restrict__ sums1, double* sums2,
           const double* __restrict__ SY, const double* __restrict__ YY, const double* __restrict__ SS,
           double* __restrict__ coef, int* __restrict__ status_dev, volatile int* status_host,
           volatile double* info_host, volatile unsigned long long* seq_host)
{
    const int m = A.msize, used = A.used;
    const int P = ka2_record(m);
    if (A.nblocks > 0) { reduce_partials(partials, A.nblocks, P, sums2); __syncthreads(); }
    bool comm_ok = true;
    if (pa.world > 1) comm_ok = p2p_allreduce_cta(pa, sums2, P);
    if (!A.do_solve) return;
    __shared__ AdaStage st;
    ada_stage_and_solve_u(A, sums1, SY, SS, YY, st);
    auto ph = [&](int i) { int s = A.oldest + i; return s >= m ? s - m : s; };
    for (int i = threadIdx.x; i < used; i += kThreads) st.w[i] = st.Rm[i][i] * st.u[i] + sums2[ph(i)];
    __syncthreads();
    if (threadIdx.x >= 32) return;
    const double tt = sums2[m];
    double U = sqrt(tt);
    bool ok = finite_d(tt) && (*status_dev != ST_COMM_TIMEOUT);
    comm_ok = comm_ok && (*status_dev != ST_COMM_TIMEOUT);
    {                                        // warp-parallel forward substitution a = R^-T w: lane i owns row i
        const unsigned full = 0xffffffffu;
        const int i = threadIdx.x;
        const bool live = i < used;
        const double inv = 1.0 / (live ? st.Rm[i][i] : 1.0);
        double t = live ? st.w[i] : 0.0, ai = 0.0;
        for (int j = 0; j < used; ++j) {
            const double aj = __shfl_sync(full, t * inv, j);
            if (i == j) ai = aj;
            if (live && i > j) t = fma(-st.Rm[j][i], aj, t);
        }
        double term = 0.0;
        bool fin = true;
        if (live) {
            coef[ph(i)] = ai;
            term = fabs(ai) * sqrt(st.ssv[i]);
            fin = finite_d(ai) && finite_d(st.u[i]);
        }
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) term += __shfl_xor_sync(full, term, o);
        U += term;
        ok = ok && __all_sync(full, fin);
    }
    if (threadIdx.x != 0) return;
    ok = ok && finite_d(U);
    coef[2 * m] = 1.0;
    coef[2 * m + 1] = U;
    coef[2 * m + 2] = tt;
    int stt = ST_ACCEPT;
    if (A.check_nan) {
        if (!ok) stt = ST_REJECT_NONFINITE;
        else if (!(U <= 0.99 * A.limit)) stt = ST_NEED_EXACT_NORM;
    }
    if (!comm_ok) stt = ST_COMM_TIMEOUT;
    *status_dev = stt;
    *status_host = stt;
    info_host[0] = U;
    info_host[1] = 1.0;
    info_host[2] = tt;
    publish_seq(seq_host, A.seq);
}

// KA3: adaQN combine + update:  h = g/sqrt(G+eps) ;  d = (used ? h*(g + sum b_j y_j) + sum a_j s_j : h)
//      MODE_AVG: x -= step*d ; x_sum += x ; grad <- d      MODE_DIRONLY: grad <- d (+ exact norm partials)
// Same work split as K3 (128-thread CTAs = 2 row-groups x 64 chunk-lanes, one shared-memory exchange per
// iteration): group 0 streams the S rows (+ x, x_sum) and forms sum a_j s_j, group 1 streams the Y rows (+ g, G)
// and forms h.(g + sum b_j y_j); group 0 then updates x and x_sum, group 1 stores grad.
template <typename T, int RPG, int MODE, int VEC>
__global__ void __launch_bounds__(k3Threads, k3_min_blocks(RPG))
ka3_combine(const T* g_in, T* grad_out, const T* __restrict__ G, const T* __restrict__ S, const T* __restrict__ Y,
            size_t ld, int msize, int used, long long n, T* __restrict__ x, T* __restrict__ x_sum, T step,
            T scal_reg, const double* __restrict__ coef, const int* __restrict__ status_dev, int force,
            double* __restrict__ partials)
{
    if (!force && *status_dev != ST_ACCEPT) return;
    const int group = threadIdx.x / k3Lanes, lane = threadIdx.x % k3Lanes;
    const T* rows[RPG];
    T cf[RPG];
    #pragma unroll
    for (int r = 0; r < RPG; ++r) {
        const bool live = r < used;
        const int j = live ? r : (used > 0 ? used - 1 : 0);
        rows[r] = (group == 0 ? S : Y) + (size_t) j * ld;
        cf[r] = live ? (T) coef[group == 0 ? j : msize + j] : (T) 0;
    }
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
            Pack<T, V> rv[RPG], gv, Gv;
            if (used > 0) {
                #pragma unroll
                for (int r = 0; r < RPG; ++r) rv[r] = ld_stream<T, V>(rows[r] + off);
            }
            if (group == 0) {
                if constexpr (MODE != MODE_DIRONLY) { xv = ld_rw<T, V>(x + off); xs = ld_rw<T, V>(x_sum + off); }
            } else {
                gv = ld_rw<T, V>(g_in + off);
                Gv = ld_stream<T, V>(G + off);
                #pragma unroll
                for (int e = 0; e < V; ++e) part.set(e, gv.get(e));
            }
            if (used > 0) {
                #pragma unroll
                for (int r = 0; r < RPG; ++r) {
                    #pragma unroll
                    for (int e = 0; e < V; ++e) part.set(e, fma(cf[r], rv[r].get(e), part.get(e)));
                }
            }
            if (group == 1) {
                #pragma unroll
                for (int e = 0; e < V; ++e) {
                    const T h = gv.get(e) / sqrt(Gv.get(e) + scal_reg);                // stochqn.c:778 / 781
                    part.set(e, used > 0 ? h * part.get(e) : h);
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
            for (int e = 0; e < V; ++e) d.set(e, xchg[buf][1][lane * V + e] + xchg[buf][0][lane * V + e]);
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
                    #pragma unroll
                    for (int e = 0; e < V; ++e) xs.set(e, xs.get(e) + xv.get(e));
                    st_vec<T, V>(x_sum + off, xs);
                } else if (grad_out) st_vec<T, V>(grad_out + off, d);
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

}  // namespace sqn
