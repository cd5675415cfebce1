// mn_small.cuh - the multinomial-logistic gradient of a SMALL batch as a whole-grid device function (see multinomial.cu for
// the conventions: scikit-learn <= 1.0 _multinomial_loss_grad, which the reference calls at stochqn/_logistic.py:7-13).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace mnsmall {

// ---- small batches: the whole gradient in ONE cooperative launch -----------------------------------------------------
// At the reference's own multinomial sizes (BibTeX: 50 samples x 1836 features x 159 classes, n = 292 083) the five
// launches above are 91 us of launch gaps around ~10 us of work.  mn_grad_small runs the same three products on one
// persistent grid (one 512-thread CTA per SM) with two grid barriers:
//   1  CTA c owns a chunk of <= 16 FEATURES: it loads its columns of X and of W once (they stay in shared memory for
//      phase 3), forms its partial Z_c = X[:, chunk] W[:, chunk]' (B x K, 4 x 8 register tiles) and stores it   -- barrier
//   2  CTA b (b < B) owns SAMPLE b: adds the G partial rows in CTA order (+ intercepts), log-sum-exp, and writes the
//      row D[b][:] = sw_b (softmax(z_b) - y_b)                                                                -- barrier
//   3  every CTA loads D (B x K) and forms its chunk of the gradient G[:, chunk] = D' X[:, chunk] + alpha W[:, chunk]; the
//      last CTA adds the intercept column (column sums of D)
// W and X are read from L2 / HBM exactly once; the only extra traffic is the G partial Z (G x B x K values).
// Restates scikit-learn (<= 1.0) _multinomial_loss_grad as mn_rows / mn_finish above (stochqn/_logistic.py:7-13).
constexpr int MS_T = 512;
constexpr int MS_CW = 16;
constexpr int MS_MAX_BK = 12288, MS_MAX_B = 128, MS_MAX_K = 512, MS_MAX_GRID = 160;


// 16-byte L2 load (ld.global.cg.v4 / v2) of VE = 16 / sizeof(T) elements
template <typename T> struct Pack16 {
    T v[16 / sizeof(T)];
    __device__ static Pack16 zero() { Pack16 p; for (int e = 0; e < (int) (16 / sizeof(T)); ++e) p.v[e] = (T) 0; return p; }
};
template <typename T> __device__ __forceinline__ Pack16<T> ld_cg16(const T* p);
template <> __device__ __forceinline__ Pack16<double> ld_cg16<double>(const double* p)
{
    const double2 t = __ldcg(reinterpret_cast<const double2*>(p));
    Pack16<double> r; r.v[0] = t.x; r.v[1] = t.y; return r;
}
template <> __device__ __forceinline__ Pack16<float> ld_cg16<float>(const float* p)
{
    const float4 t = __ldcg(reinterpret_cast<const float4*>(p));
    Pack16<float> r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r;
}

// Grid barrier of the stand-alone kernel: its counter is zeroed by the host before every launch, so the k-th barrier is passed when
// the counter reaches k * gridDim.x - a fire-and-forget `red` to arrive, no ticket round trip (as RunBarrier, kernels_loop.cuh).
struct MsBarrier {
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

template <typename T>
struct MnSmallArgs {
    const T* X; long long ldx;
    const T* Y; long long ldy;
    const int* labels;
    const T* sw;
    int B, d, K, icpt;
    const T* W;
    T alpha;
    T* Gout;
    T* Zp;                      // [gridDim.x][B][KZ] partial products
    T* Dg;                      // [B][K]
};

// The body: a whole-grid function (every CTA of a cooperative grid calls it with the same arguments).  `ms_smem` = the dynamic
// shared memory (mn_small_smem bytes); `barrier()` = a grid barrier.  Used by the stand-alone kernel below (multinomial.cu) and by
// the persistent adaQN request loop (kernels_loop.cuh: kl_fit_mn_ada).
template <typename T, typename Barrier>
__device__ __forceinline__ void mn_grad_small_body(const MnSmallArgs<T>& a, unsigned char* ms_smem, Barrier&& barrier, unsigned long long* trace)
{
    const T* __restrict__ X = a.X; const long long ldx = a.ldx; const T* __restrict__ Y = a.Y; const long long ldy = a.ldy;
    const int* __restrict__ labels = a.labels; const T* __restrict__ sw = a.sw;
    const int B = a.B, d = a.d, K = a.K, icpt = a.icpt;
    const T* W = a.W; const T alpha = a.alpha; T* __restrict__ Gout = a.Gout; T* Zp = a.Zp; T* Dg = a.Dg;
    const int G = (int) gridDim.x, c = (int) blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cw = (d + G - 1) / G;
    const int j0 = c * cw < d ? c * cw : d;
    const int j1 = j0 + cw < d ? j0 + cw : d;
    const int w = j1 - j0;
    const long long ldw = (long long) d + (icpt ? 1 : 0);
    T* Xc = reinterpret_cast<T*>(ms_smem);                   // [cw][B]   (sample index contiguous)
    T* Wc = Xc + (size_t) cw * B;                            // [cw][K]   (class index contiguous)
    T* Ds = reinterpret_cast<T*>((reinterpret_cast<uintptr_t>(Wc + (size_t) cw * K) + 15) & ~(uintptr_t) 15);      // [B][K]
    __shared__ double red_s[MS_T / 32];
    __shared__ double bc_s[2];
    auto stamp = [&](int k) {                                // development aid (stochqn_b200_debug_mn_trace)
        if (trace && c == 0 && tid == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); trace[k] = t; }
    };
    stamp(0);

    // ---- phase 1: the chunk of X and W, partial Z ----
    constexpr int VE = 16 / (int) sizeof(T);                 // elements of a 16-byte vector
    const int KZ = (K + VE - 1) / VE * VE;                   // row pitch of the partial Z records (16-byte aligned rows)
    // (all the loads of a thread are issued before its first shared-memory store: a load -> store pair per trip would cost one
    // memory round trip each - five for the W chunk at the config-3 shape)
    for (int t0 = tid; t0 < B * cw; t0 += 4 * MS_T) {
        T v[4];
        #pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int t = t0 + u * MS_T;
            const int b = t / cw, jj = t % cw;
            v[u] = (t < B * cw && jj < w) ? __ldg(X + (long long) b * ldx + j0 + jj) : (T) 0;
        }
        #pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int t = t0 + u * MS_T;
            if (t < B * cw) Xc[(t % cw) * B + t / cw] = v[u];
        }
    }
    for (int t0 = tid; t0 < K * cw; t0 += 6 * MS_T) {
        T v[6];
        #pragma unroll
        for (int u = 0; u < 6; ++u) {
            const int t = t0 + u * MS_T;
            const int k = t / cw, jj = t % cw;
            v[u] = (t < K * cw && jj < w) ? __ldcg(W + (long long) k * ldw + j0 + jj) : (T) 0;
        }
        #pragma unroll
        for (int u = 0; u < 6; ++u) {
            const int t = t0 + u * MS_T;
            if (t < K * cw) Wc[(t % cw) * K + t / cw] = v[u];
        }
    }
    __syncthreads();
    {
        const int BT = (B + 1) / 2, KT = (K + 7) / 8;        // thread tile: b in {bt, bt + BT}, k in {kt + KT*ik}: lanes walk kt -> conflict-free
        T* zout = Zp + (size_t) c * (size_t) B * KZ;
        for (int t = tid; t < BT * KT; t += MS_T) {
            const int bt = t / KT, kt = t % KT;
            const int b0 = bt, b1 = bt + BT;
            T acc[2][8];
            #pragma unroll
            for (int ik = 0; ik < 8; ++ik) { acc[0][ik] = (T) 0; acc[1][ik] = (T) 0; }
            for (int jj = 0; jj < w; ++jj) {
                T wa[8];
                const T x0 = Xc[jj * B + b0], x1 = b1 < B ? Xc[jj * B + b1] : (T) 0;
                #pragma unroll
                for (int ik = 0; ik < 8; ++ik) { const int k = kt + KT * ik; wa[ik] = k < K ? Wc[jj * K + k] : (T) 0; }
                #pragma unroll
                for (int ik = 0; ik < 8; ++ik) { acc[0][ik] = fma(x0, wa[ik], acc[0][ik]); acc[1][ik] = fma(x1, wa[ik], acc[1][ik]); }
            }
            #pragma unroll
            for (int ik = 0; ik < 8; ++ik) {
                const int k = kt + KT * ik;
                if (k < K) {
                    zout[(size_t) b0 * KZ + k] = acc[0][ik];
                    if (b1 < B) zout[(size_t) b1 * KZ + k] = acc[1][ik];
                }
            }
        }
        if (KZ > K) {                                        // the padding of every row is read (as part of a vector) by phase 2
            for (int t = tid; t < B * (KZ - K); t += MS_T) zout[(size_t) (t / (KZ - K)) * KZ + K + t % (KZ - K)] = (T) 0;
        }
    }
    stamp(1);
    barrier();
    stamp(2);

    // ---- phase 2: one CTA per sample: sum the partial rows (16-byte loads, 8 records in flight per thread), softmax, D row ----
    {
        const int NV = KZ / VE;                              // vectors per row
        const int NVP = (NV + 31) / 32 * 32;
        const int NQ = MS_T / NVP;                           // record slices (K <= 512: at least one)
        double* zrow = reinterpret_cast<double*>(Ds);         // (Ds is not in use yet) [NQ][NVP * VE] partial sums
        const size_t rs = (size_t) B * KZ;
        for (int b = c; b < B; b += G) {
            const int v = tid % NVP, q = tid / NVP;
            if (q < NQ && v < NV) {
                double a[VE];
                #pragma unroll
                for (int e = 0; e < VE; ++e) a[e] = 0.0;
                const T* pz = Zp + (size_t) b * KZ + (size_t) v * VE;
                for (int r0 = q; r0 < G; r0 += 8 * NQ) {
                    Pack16<T> pv[8];
                    #pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int r = r0 + u * NQ;
                        if (r < G) pv[u] = ld_cg16<T>(pz + (size_t) r * rs); else pv[u] = Pack16<T>::zero();
                    }
                    #pragma unroll
                    for (int u = 0; u < 8; ++u)
                        #pragma unroll
                        for (int e = 0; e < VE; ++e) a[e] += (double) pv[u].v[e];
                }
                #pragma unroll
                for (int e = 0; e < VE; ++e) zrow[(size_t) q * NVP * VE + v * VE + e] = a[e];
            }
            __syncthreads();
            double z = -INFINITY;
            if (tid < K) {
                z = icpt ? (double) __ldcg(W + (long long) tid * ldw + d) : 0.0;
                for (int q2 = 0; q2 < NQ; ++q2) z += zrow[(size_t) q2 * NVP * VE + tid];
            }
            double mx = z;
            for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            if (lane == 0) red_s[warp] = mx;
            __syncthreads();
            if (tid == 0) { double v2 = red_s[0]; for (int q2 = 1; q2 < MS_T / 32; ++q2) v2 = fmax(v2, red_s[q2]); bc_s[0] = v2; }
            __syncthreads();
            mx = bc_s[0];
            const double ez = tid < K ? exp(z - mx) : 0.0;
            double se = ez;
            for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
            if (lane == 0) red_s[warp] = se;
            __syncthreads();
            if (tid == 0) { double v2 = 0; for (int q2 = 0; q2 < MS_T / 32; ++q2) v2 += red_s[q2]; bc_s[1] = v2; }
            __syncthreads();
            if (tid < K) {
                const double lse = mx + log(bc_s[1]);
                const double pk = exp(z - lse);                                  // as mn_rows: exp(z - lse)
                const double wt = sw ? (double) sw[b] : 1.0;
                const double yk = labels ? (labels[b] == tid ? 1.0 : 0.0) : (double) Y[(long long) b * ldy + tid];
                Dg[(size_t) b * K + tid] = (T) (wt * (pk - yk));
            }
            __syncthreads();
        }
    }
    stamp(3);
    barrier();
    stamp(4);

    // ---- phase 3: G[:, chunk] = D' X[:, chunk] + alpha W[:, chunk]; the intercept column by the last CTA ----
    {
        const int total = B * K;
        for (int t0 = tid; t0 < total; t0 += 8 * MS_T) {     // eight loads in flight per thread
            T dv[8];
            #pragma unroll
            for (int u = 0; u < 8; ++u) { const int t = t0 + u * MS_T; dv[u] = t < total ? __ldcg(Dg + t) : (T) 0; }
            #pragma unroll
            for (int u = 0; u < 8; ++u) { const int t = t0 + u * MS_T; if (t < total) Ds[t] = dv[u]; }
        }
    }
    __syncthreads();
    stamp(5);
    {
        const int KP = (K + 31) / 32 * 32;
        const int NH = MS_T / KP;                            // feature interleave: thread (k, h) owns jj = h, h + NH, ... (h is warp-uniform)
        const int k = tid % KP, h = tid / KP;
        if (k < K && h < NH) {
            #pragma unroll
            for (int ug = 0; ug < MS_CW / 4; ++ug) {         // four features at a time; groups beyond the chunk are skipped (warp-uniform)
                const int jb = h + 4 * ug * NH;
                if (jb < w) {
                    const int j1 = jb + NH, j2 = jb + 2 * NH, j3 = jb + 3 * NH;
                    const T* x0 = Xc + jb * B;
                    const T* x1 = Xc + (j1 < w ? j1 : jb) * B;
                    const T* x2 = Xc + (j2 < w ? j2 : jb) * B;
                    const T* x3 = Xc + (j3 < w ? j3 : jb) * B;
                    T a0 = (T) 0, a1 = (T) 0, a2 = (T) 0, a3 = (T) 0;
                    #pragma unroll 5
                    for (int b = 0; b < B; ++b) {
                        const T dv = Ds[b * K + k];
                        a0 = fma(dv, x0[b], a0); a1 = fma(dv, x1[b], a1); a2 = fma(dv, x2[b], a2); a3 = fma(dv, x3[b], a3);
                    }
                    T* go = Gout + (long long) k * ldw + j0;
                    go[jb] = fma(alpha, Wc[jb * K + k], a0);
                    if (j1 < w) go[j1] = fma(alpha, Wc[j1 * K + k], a1);
                    if (j2 < w) go[j2] = fma(alpha, Wc[j2 * K + k], a2);
                    if (j3 < w) go[j3] = fma(alpha, Wc[j3 * K + k], a3);
                }
            }
        }
        if (icpt && c == G - 1 && tid < K) {
            double sacc = 0.0;
            for (int b = 0; b < B; ++b) sacc += (double) Ds[b * K + tid];
            Gout[(long long) tid * ldw + d] = (T) sacc;
        }
    }
    if (trace) { __syncthreads(); stamp(6); }
}


template <typename T>
__global__ void __launch_bounds__(MS_T, 1)
mn_grad_small(const MnSmallArgs<T> a, unsigned long long* bar, unsigned long long* trace)
{
    extern __shared__ __align__(16) unsigned char ms_dyn_smem[];
    MsBarrier barrier{bar, 0ull};
    mn_grad_small_body<T>(a, ms_dyn_smem, barrier, trace);
}

}  // namespace mnsmall
