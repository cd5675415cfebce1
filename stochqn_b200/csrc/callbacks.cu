// callbacks.cu - bundled device callbacks that serve the optimizers' requests without leaving
// the GPU: the chained Rosenbrock function of the reference's example programs
// (example/c_rosen.c:13-41) and the binary-logistic closed forms of the reference's R model
// (R/logistic.R:1-37).  C ABI declared in include/stochqn_b200.h.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <atomic>

#include "stochqn.h"
#include "stochqn_b200.h"
#include "vecio.cuh"
#include "p2p.cuh"
#include "logistic_form.cuh"

// stochqn_b200.cu: PeerArgs of the next exchange on a communicator (world = 0 when the exchange must go through NCCL)
sqn::PeerArgs stochqn_b200_internal_next_exchange(void* comm, size_t count, int** error_flag, cudaStream_t stream);

// kernels launched by the callbacks; added to stochqn_b200_launch_count() by stochqn_b200.cu
std::atomic<unsigned long long> stochqn_b200_cb_launches{0};

namespace {

constexpr int kT = 256;

int check_launch(const char* what, int launched = 1)
{
    stochqn_b200_cb_launches.fetch_add((unsigned long long) launched, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        fprintf(stderr, "stochqn_b200: %s launch failed: %s\n", what, cudaGetErrorString(e));
        return -2;
    }
    return 0;
}

int grid_1d(long long items, int cap = 148 * 8)
{
    long long g = (items + kT - 1) / kT;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int) g;
}

// ---- Rosenbrock ---------------------------------------------------------------------------------
// x0[i] = 0.95 + 1e-4 * ((uint32)(i * 2654435761) mod 1000): pure integer hash, identical in C / NumPy / CUDA
__global__ void rosen_x0_kernel(real_t* __restrict__ x, long long n, long long offset)
{
    const long long stride = (long long) gridDim.x * blockDim.x;
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t h = (uint32_t) ((uint64_t) (i + offset) * 2654435761ull);
        x[i] = (real_t) __dadd_rn(0.95, __dmul_rn(1e-4, (double) (h % 1000u)));     // no FMA contraction: bit-identical to C / NumPy
    }
}

// interior: 200(x_i - x_{i-1}^2) - 400(x_{i+1} - x_i^2) x_i - 2(1 - x_i)      (c_rosen.c:35-40)
// first   : -400 x_0 (x_1 - x_0^2) - 2(1 - x_0)                                (c_rosen.c:28)
// last    : 200 (x_{n-1} - x_{n-2}^2)                                          (c_rosen.c:29)
__device__ __forceinline__ double rosen_g(double xm, double xc, double xp, bool has_left, bool has_right)
{
    double out = 0.0;
    if (has_left) out += 200.0 * (xc - xm * xm);
    if (has_right) { out -= 400.0 * (xp - xc * xc) * xc; out -= 2.0 * (1.0 - xc); }
    return out;
}

// One 16-byte chunk per thread and iteration: the chunk is loaded as a vector; its two neighbours are the edge
// elements of the adjacent lanes' chunks and arrive by warp shuffle (only lanes 0 and 31 load a scalar), the
// result is stored as a vector.
// SHARDED: the halo exchange is fused into this kernel.  The chunks that touch the ends of the shard are left out
// of the streaming loop; CTA 0 exchanges (first, last) of every rank over the peer-memory mailboxes (p2p.cuh)
// while the other CTAs stream, then computes those few edge elements.
template <int VEC, bool SHARDED, int U, bool SHUFFLE>
__global__ void __launch_bounds__(kT)
rosen_grad_kernel(const real_t* __restrict__ x, real_t* __restrict__ g, long long n, long long offset,
                  long long n_global, const real_t* __restrict__ halo, sqn::PeerArgs pa, int* __restrict__ error_flag,
                  real_t* __restrict__ halo_out)
{
    const long long nchunks = n / VEC;
    const long long stride = (long long) gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    for (long long base = (long long) blockIdx.x * blockDim.x + (threadIdx.x & ~31); base < nchunks; base += U * stride) {   // warp-uniform
        sqn::Pack<real_t, VEC> xv[U];
        double le[U], re[U];
        bool valid[U];
        // all loads of the iteration first: the vector chunk, and for the two end lanes of the warp the scalar neighbour
        #pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long c = base + u * stride + lane;
            const long long i0 = c * VEC;
            valid[u] = c < nchunks;
            le[u] = 0.0; re[u] = 0.0;
            if (valid[u]) {
                xv[u] = sqn::ld_stream<real_t, VEC>(x + i0);
                if (!SHUFFLE || lane == 0) le[u] = (i0 > 0) ? (double) __ldg(x + i0 - 1) : ((halo && offset > 0) ? (double) halo[0] : 0.0);
                if (!SHUFFLE || lane == 31 || c == nchunks - 1)
                    re[u] = (i0 + VEC < n) ? (double) __ldg(x + i0 + VEC) : ((halo && offset + n < n_global) ? (double) halo[1] : 0.0);
            } else {
                #pragma unroll
                for (int e = 0; e < VEC; ++e) xv[u].set(e, (real_t) 0);
            }
        }
        #pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long c = base + u * stride + lane;
            const long long i0 = c * VEC;
            double v[VEC + 2];
            #pragma unroll
            for (int e = 0; e < VEC; ++e) v[e + 1] = (double) xv[u].get(e);
            if (SHUFFLE) {
                v[0] = __shfl_up_sync(0xffffffffu, v[VEC], 1);
                v[VEC + 1] = __shfl_down_sync(0xffffffffu, v[1], 1);
            }
            if (!valid[u]) continue;
            if (SHARDED && (c == 0 || c == nchunks - 1)) continue;          // shard ends: CTA 0, after the exchange
            if (!SHUFFLE || lane == 0) v[0] = le[u];
            if (!SHUFFLE || lane == 31 || c == nchunks - 1) v[VEC + 1] = re[u];
            sqn::Pack<real_t, VEC> gv;
            #pragma unroll
            for (int e = 0; e < VEC; ++e) {
                const long long gi = i0 + e + offset;
                gv.set(e, (real_t) rosen_g(v[e], v[e + 1], v[e + 2], gi > 0, gi < n_global - 1));
            }
            sqn::st_vec<real_t, VEC>(g + i0, gv);
        }
    }
    if (blockIdx.x != 0) return;
    if (!SHARDED) {
        if (VEC > 1) {                                  // scalar tail
            const long long i = nchunks * VEC + threadIdx.x;
            if (i < n) {
                const long long gi = i + offset;
                const double xm = (i > 0) ? (double) x[i - 1] : ((halo && offset > 0) ? (double) halo[0] : 0.0);
                const double xp = (i < n - 1) ? (double) x[i + 1] : ((halo && offset + n < n_global) ? (double) halo[1] : 0.0);
                g[i] = (real_t) rosen_g(xm, (double) x[i], xp, gi > 0, gi < n_global - 1);
            }
        }
        return;
    }
    // ---- sharded: exchange (first, last) of every rank, then the edge elements ----
    __shared__ double rec[2 * sqn::kMaxWorld];
    const int t = threadIdx.x;
    if (t < 2 * pa.world) rec[t] = (t == 2 * pa.rank) ? (double) x[0] : (t == 2 * pa.rank + 1) ? (double) x[n - 1] : 0.0;
    if (!sqn::p2p_allreduce_cta(pa, rec, 2 * pa.world) && t == 0 && error_flag) *error_flag = 1;
    const double hl = pa.rank > 0 ? rec[2 * (pa.rank - 1) + 1] : 0.0;
    const double hr = pa.rank < pa.world - 1 ? rec[2 * (pa.rank + 1)] : 0.0;
    if (t == 0 && halo_out) { halo_out[0] = (real_t) hl; halo_out[1] = (real_t) hr; }   // kept for re-evaluations at the same point
    const long long a_end = n < VEC ? n : VEC;                           // [0, a_end): first chunk
    long long b_lo = (nchunks - 1) * VEC;                                // [b_lo, n): last chunk + scalar tail
    if (b_lo < a_end) b_lo = a_end;
    for (int pass = 0; pass < 2; ++pass) {
        const long long i = pass == 0 ? (long long) t : b_lo + t;
        const long long hi = pass == 0 ? a_end : n;
        if (i < hi) {
            const long long gi = i + offset;
            const double xm = (i > 0) ? (double) x[i - 1] : hl;
            const double xp = (i < n - 1) ? (double) x[i + 1] : hr;
            g[i] = (real_t) rosen_g(xm, (double) x[i], xp, gi > 0, gi < n_global - 1);
        }
    }
}

// halo exchange for sharded Rosenbrock: every rank contributes its first and last element to a zero-padded
// [2*world] record, one sum all-reduce turns it into an all-gather, then each rank picks its neighbours' values
__global__ void rosen_halo_pack(const real_t* __restrict__ x, long long n, int rank, int world, double* __restrict__ rec)
{
    const int t = threadIdx.x;
    if (t < 2 * world) rec[t] = (t == 2 * rank) ? (double) x[0] : (t == 2 * rank + 1) ? (double) x[n - 1] : 0.0;
}
__global__ void rosen_halo_unpack(const double* __restrict__ rec, int rank, int world, real_t* __restrict__ halo)
{
    if (threadIdx.x == 0) halo[0] = rank > 0 ? (real_t) rec[2 * (rank - 1) + 1] : (real_t) 0;
    if (threadIdx.x == 1) halo[1] = rank < world - 1 ? (real_t) rec[2 * (rank + 1)] : (real_t) 0;
}

__device__ double g_fun_partials[2048];
__device__ unsigned int g_fun_ticket = 0;

// f = sum_{i < n_global-1} 100 (x_{i+1} - x_i^2)^2 + (1 - x_i)^2               (c_rosen.c:13-24)
// deterministic: CTA partials, the last CTA to finish adds them in index order
__global__ void __launch_bounds__(kT)
rosen_fun_kernel(const real_t* __restrict__ x, long long n, long long offset, long long n_global,
                 const real_t* __restrict__ halo, double* __restrict__ f_out)
{
    double acc = 0.0;
    const long long stride = (long long) gridDim.x * blockDim.x;
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (i + offset < n_global - 1) {
            const double xc = (double) x[i];
            const double xp = (i < n - 1) ? (double) x[i + 1] : (double) halo[1];
            const double d1 = xp - xc * xc, d2 = 1.0 - xc;
            acc += 100.0 * d1 * d1 + d2 * d2;
        }
    }
    __shared__ double red[kT / 32];
    __shared__ bool last;
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double v = 0;
        for (int w = 0; w < kT / 32; ++w) v += red[w];
        g_fun_partials[blockIdx.x] = v;
        __threadfence();
        last = (atomicAdd(&g_fun_ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        double v = 0;
        for (unsigned b = 0; b < gridDim.x; ++b) v += ((volatile double*) g_fun_partials)[b];
        *f_out = v;
        g_fun_ticket = 0;
    }
}

// ---- binary logistic regression: loss, and the two-sweep gradient / Hessian-vector form kept for very wide
// ---- matrices (ncols > 5120); the fused one-sweep kernel is further down ------------------------------------
// pass A: one warp per row: z = x_row'w (and t = x_row'v), r_row written to scratch
//         grad: r = (sigmoid(z) - y) * sw          hess_vec: r = p(1-p) * sw * t        loss: per-row loss * sw
// pass B: one thread per column: out[col] = sum_rows r_row X[row][col] / sum(sw) + 2*lambda*u[col]
template <int KIND>
__global__ void __launch_bounds__(kT)
logistic_rows_kernel(const real_t* __restrict__ X, long long ldx, const real_t* __restrict__ y,
                     const real_t* __restrict__ sw, long long nrows, long long ncols,
                     const real_t* __restrict__ w, const real_t* __restrict__ v, double* __restrict__ r, const LgForm form)
{
    const int lane = threadIdx.x & 31;
    const double zc = form.icpt ? (double) w[ncols] : 0.0;
    const double tc = (form.icpt && KIND == LG_HVP) ? (double) v[ncols] : 0.0;
    const long long warp = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long) gridDim.x * blockDim.x) >> 5;
    for (long long row = warp; row < nrows; row += nwarps) {
        const real_t* xr = X + row * ldx;
        double z = 0, t = 0;
        for (long long c = lane; c < ncols; c += 32) {
            const double xv = (double) xr[c];
            z = fma(xv, (double) w[c], z);
            if (KIND == LG_HVP) t = fma(xv, (double) v[c], t);
        }
        for (int o = 16; o > 0; o >>= 1) {
            z += __shfl_down_sync(0xffffffffu, z, o);
            if (KIND == LG_HVP) t += __shfl_down_sync(0xffffffffu, t, o);
        }
        if (lane == 0) {
            const double wt = sw ? (double) sw[row] : 1.0;
            r[row] = lg_row_weight(KIND, form, z + zc, t + tc, (double) y[row], wt);
        }
    }
}

// r[nrows] -> scalars: r[nrows] = sum(sw) (or nrows), r[nrows+1] = sum(r) ; single CTA, fixed order
__global__ void __launch_bounds__(kT)
logistic_norm_kernel(const real_t* __restrict__ sw, long long nrows, double* __restrict__ r)
{
    double a = 0, b = 0;
    for (long long i = threadIdx.x; i < nrows; i += kT) { a += sw ? (double) sw[i] : 1.0; b += r[i]; }
    __shared__ double ra[kT / 32], rb[kT / 32];
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_down_sync(0xffffffffu, a, o); b += __shfl_down_sync(0xffffffffu, b, o); }
    if ((threadIdx.x & 31) == 0) { ra[threadIdx.x >> 5] = a; rb[threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double sa = 0, sb = 0;
        for (int k = 0; k < kT / 32; ++k) { sa += ra[k]; sb += rb[k]; }
        r[nrows] = sa;
        r[nrows + 1] = sb;
    }
}

// columns: out[c] = (sum_rows r[row] X[row][c]) / r[nrows] + 2*lambda*u[c]; rows split over blockIdx.y with
// per-slice partial columns summed in fixed order by the finishing pass
__global__ void __launch_bounds__(kT)
logistic_cols_kernel(const real_t* __restrict__ X, long long ldx, long long nrows, long long ncols,
                     const double* __restrict__ r, double* __restrict__ colpart)
{
    const long long c = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncols) return;
    const long long per = (nrows + gridDim.y - 1) / gridDim.y;
    const long long r0 = (long long) blockIdx.y * per;
    const long long r1 = r0 + per < nrows ? r0 + per : nrows;
    double acc = 0;
    for (long long row = r0; row < r1; ++row) acc = fma(r[row], (double) X[row * ldx + c], acc);
    colpart[(long long) blockIdx.y * ncols + c] = acc;
}

__global__ void __launch_bounds__(kT)
logistic_finish_kernel(const double* __restrict__ colpart, int slices, long long nrows, long long ncols,
                       const double* __restrict__ r, const real_t* __restrict__ u, real_t lambda,
                       real_t* __restrict__ out, const LgForm form)
{
    const long long c = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (c == ncols && form.icpt) { out[c] = (real_t) r[nrows + 1]; return; }       // intercept: sum of the row weights
    if (c >= ncols) return;
    double acc = 0;
    for (int s = 0; s < slices; ++s) acc += colpart[(long long) s * ncols + c];
    out[c] = form.sk ? (real_t) (acc + (double) lambda * (double) u[c])
                     : (real_t) (acc / r[nrows] + 2.0 * (double) lambda * (double) u[c]);
}

__global__ void logistic_loss_finish_kernel(const double* __restrict__ r, long long nrows, const real_t* __restrict__ w,
                                            long long ncols, real_t lambda, double* __restrict__ loss, const LgForm form)
{
    double a = 0;
    for (long long i = threadIdx.x; i < ncols; i += blockDim.x) { const double wv = (double) w[i]; a = fma(wv, wv, a); }
    __shared__ double ra[32];
    for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) ra[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0;
        for (unsigned k = 0; k < blockDim.x / 32; ++k) s += ra[k];
        *loss = form.sk ? r[nrows + 1] + 0.5 * (double) lambda * s : r[nrows + 1] / r[nrows] + (double) lambda * s;
    }
}

// ---- fused gradient / Hessian-vector kernel: ONE sweep of the batch --------------------------------------------
// The first version above reads the batch twice (row pass for z = Xw, column pass for X'r).  Here a 256-thread CTA
// owns a strided set of rows and keeps, spread over its threads' registers, a full-width fp64 column accumulator
// (thread t owns columns t, t+256, ...: CPT of them).  Per iteration it loads R whole rows (thread t its CPT columns
// of each: coalesced 8/4-byte loads, all issued before the first use), reduces the R dot products x_i'w (and x_i'v)
// across the CTA (warp shuffles + one barrier, double-buffered), turns them into the row weights r_i and adds
// r_i * x_i into the accumulator from the registers that still hold the row - X is read from HBM exactly once.
// CTA partial columns (+ the CTA's sum of sample weights) go to `colpart`; logistic_fused_finish adds them in CTA
// order (deterministic), divides by sum(sw) and adds 2*lambda*u   (R/logistic.R:12-21, 23-37).
constexpr int LT = 256;
__host__ __device__ constexpr int lg_rows(int cpt) { return cpt <= 4 ? 8 : cpt <= 8 ? 4 : 2; }
constexpr int kLgMaxGrid = 320;          // CTA partial records (two CTAs per SM)
constexpr int kLgMaxCpt = 20;            // columns per thread held in registers: ncols <= 5120 takes the fused path

// products and the per-thread partial sums run in the storage type (fp64 for the double build - the reference's R
// arithmetic; fp32 FMA for the float build, where an F2F conversion per element would be the bottleneck); everything
// that crosses threads or rows of different CTAs is fp64.
template <int CPT, int KIND, int R, int MINB>
__global__ void __launch_bounds__(LT, MINB)
logistic_fused_kernel(const real_t* __restrict__ X, long long ldx, const real_t* __restrict__ y,
                      const real_t* __restrict__ sw, long long nrows, long long ncols,
                      const real_t* __restrict__ w, const real_t* __restrict__ v, double* __restrict__ colpart,
                      const LgForm form)
{
    extern __shared__ __align__(16) unsigned char lg_smem[];
    real_t* ws = reinterpret_cast<real_t*>(lg_smem);              // w (and v) staged once per CTA: [CPT*LT] each
    real_t* vs = ws + CPT * LT;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    real_t acc[CPT];
    #pragma unroll
    for (int k = 0; k < CPT; ++k) {
        const long long c = tid + (long long) k * LT;
        ws[tid + k * LT] = c < ncols ? w[c] : (real_t) 0;
        if (KIND == LG_HVP) vs[tid + k * LT] = c < ncols ? v[c] : (real_t) 0;
        acc[k] = (real_t) 0;
    }
    double sw_sum = 0.0, r_sum = 0.0;
    const double zc = form.icpt ? (double) w[ncols] : 0.0;
    const double tc = (form.icpt && KIND == LG_HVP) ? (double) v[ncols] : 0.0;
    __shared__ double red[2][2][LT / 32][R];
    __shared__ double rw[2][R], sc_s[2][R];
    int par = 0;
    // (each thread reads back only the w / v entries it wrote itself: no barrier needed before the loop)
    for (long long row0 = (long long) blockIdx.x * R; row0 < nrows; row0 += (long long) gridDim.x * R) {
        real_t xv[R][CPT];
        #pragma unroll
        for (int r = 0; r < R; ++r) {
            const long long row = row0 + r;
            const real_t* xr = X + row * ldx;
            #pragma unroll
            for (int k = 0; k < CPT; ++k) {
                const long long c = tid + (long long) k * LT;
                xv[r][k] = (row < nrows && c < ncols) ? __ldg(xr + c) : (real_t) 0;
            }
        }
        real_t zp[R], tp[R];
        #pragma unroll
        for (int r = 0; r < R; ++r) { zp[r] = (real_t) 0; tp[r] = (real_t) 0; }
        #pragma unroll
        for (int k = 0; k < CPT; ++k) {
            const real_t wk = ws[tid + k * LT];
            real_t vk = (real_t) 0;
            if (KIND == LG_HVP) vk = vs[tid + k * LT];
            #pragma unroll
            for (int r = 0; r < R; ++r) {
                zp[r] = fma(xv[r][k], wk, zp[r]);
                if (KIND == LG_HVP) tp[r] = fma(xv[r][k], vk, tp[r]);
            }
        }
        #pragma unroll
        for (int r = 0; r < R; ++r) {
            double z = (double) zp[r], t = (double) tp[r];
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                z += __shfl_down_sync(0xffffffffu, z, o);
                if (KIND == LG_HVP) t += __shfl_down_sync(0xffffffffu, t, o);
            }
            if (lane == 0) { red[par][0][warp][r] = z; if (KIND == LG_HVP) red[par][1][warp][r] = t; }
        }
        __syncthreads();
        // the row weights once per row (thread r < R), not once per thread: the fp64 exp + divide are ~100 issue slots, and
        // 256 threads each doing all R of them cost as much as streaming a 1000-column row
        if (tid < R) {
            const long long row = row0 + tid;
            double z = 0, t = 0;
            #pragma unroll
            for (int q = 0; q < LT / 32; ++q) { z += red[par][0][q][tid]; if (KIND == LG_HVP) t += red[par][1][q][tid]; }
            double rr = 0.0;
            if (row < nrows) {
                const double wt = sw ? (double) sw[row] : 1.0;
                rr = lg_row_weight(KIND, form, z + zc, t + tc, (double) y[row], wt);
                sw_sum += wt; r_sum += rr;
            }
            rw[par][tid] = rr;
        }
        __syncthreads();
        #pragma unroll
        for (int r = 0; r < R; ++r) {
            const real_t rt = (real_t) rw[par][r];
            #pragma unroll
            for (int k = 0; k < CPT; ++k) acc[k] = fma(rt, xv[r][k], acc[k]);
        }
        par ^= 1;
    }
    double* out = colpart + (size_t) blockIdx.x * (size_t) (ncols + 2);
    #pragma unroll
    for (int k = 0; k < CPT; ++k) {
        const long long c = tid + (long long) k * LT;
        if (c < ncols) out[c] = (double) acc[k];
    }
    if (tid < R) { sc_s[0][tid] = sw_sum; sc_s[1][tid] = r_sum; }
    __syncthreads();
    if (tid == 0) {
        double a = 0, b = 0;
        #pragma unroll
        for (int r = 0; r < R; ++r) { a += sc_s[0][r]; b += sc_s[1][r]; }
        out[ncols] = a; out[ncols + 1] = b;
    }
}

// 32 columns x 8 record-slices per CTA: thread (tx, ty) adds the records ty, ty+8, ... of column tx (CTA order within
// the slice), the 8 slice sums are added in slice order - deterministic, and 8x shorter dependent chains than one thread
// per column walking all the records (which made this pass as long as the sweep itself for small batches).
constexpr int kFinCols = 32, kFinSlices = 8;
__global__ void __launch_bounds__(kFinCols * kFinSlices)
logistic_fused_finish(const double* __restrict__ colpart, int nparts, long long ncols, const real_t* __restrict__ u,
                      real_t lambda, real_t* __restrict__ out, const LgForm form)
{
    __shared__ double part[kFinSlices][kFinCols + 1], swp[kFinSlices], rsp[kFinSlices];
    const int tx = threadIdx.x % kFinCols, ty = threadIdx.x / kFinCols;
    const long long c = (long long) blockIdx.x * kFinCols + tx;
    const size_t rec = (size_t) (ncols + 2);
    double acc = 0;
    if (c < ncols)
        for (int b = ty; b < nparts; b += kFinSlices) acc += __ldcg(colpart + (size_t) b * rec + c);
    part[ty][tx] = acc;
    if (tx == 0) {                                          // the two scalars of every record: sum of weights, sum of row weights
        double sw = 0, rs = 0;
        for (int b = ty; b < nparts; b += kFinSlices) { sw += __ldcg(colpart + (size_t) b * rec + ncols); rs += __ldcg(colpart + (size_t) b * rec + ncols + 1); }
        swp[ty] = sw; rsp[ty] = rs;
    }
    __syncthreads();
    if (ty != 0) return;
    double swt = 0, rst = 0, tot = 0;
    #pragma unroll
    for (int q = 0; q < kFinSlices; ++q) { tot += part[q][tx]; swt += swp[q]; rst += rsp[q]; }
    if (c < ncols)
        out[c] = form.sk ? (real_t) (tot + (double) lambda * (double) u[c])
                         : (real_t) (tot / swt + 2.0 * (double) lambda * (double) u[c]);
    if (form.icpt && blockIdx.x == 0 && tx == 0) out[ncols] = (real_t) rst;      // intercept: sum of the row weights
}


// ---- CSR rows -> dense rows (sparse model matrices for the bundled callbacks) ---------------------------------------
// The reference's Python estimator keeps a scipy CSR matrix and hands it to scikit-learn's sparse-aware arithmetic
// (stochqn/_logistic.py:155).  The device callbacks are dense one-sweep kernels, so a CSR model matrix is expanded once, on the
// device, into the resident dense matrix they stream (stochqn_b200/logistic.py): one warp per row zero-fills it and scatters
// the stored entries (canonical CSR: no duplicate column in a row).
__global__ void __launch_bounds__(kT)
csr_rows_to_dense_kernel(const long long* __restrict__ indptr, const long long* __restrict__ indices, const real_t* __restrict__ data,
                         long long row0, long long nrows, long long ncols, real_t* __restrict__ out, long long ldo, int* __restrict__ bad)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long) gridDim.x * blockDim.x) >> 5;
    for (long long r = warp; r < nrows; r += nwarps) {
        real_t* o = out + r * ldo;
        for (long long j = lane; j < ncols; j += 32) o[j] = (real_t) 0;
        __syncwarp();
        const long long p0 = indptr[row0 + r], p1 = indptr[row0 + r + 1];
        for (long long p = p0 + lane; p < p1; p += 32) {
            const long long j = indices[p];
            if (j >= 0 && j < ncols) o[j] = data[p];
            else if (bad) *bad = 1;
        }
    }
}

template <int C, int KIND, int R, int MINB>
void launch_logistic_fused_t(const real_t* X, long long ldx, const real_t* y, const real_t* sw, long long nrows, long long ncols,
                             const real_t* w, const real_t* v, double* colpart, int* nparts, int sms, cudaStream_t st, LgForm form)
{
    auto kern = logistic_fused_kernel<C, KIND, R, MINB>;
    const size_t smem = (size_t) (KIND == LG_HVP ? 2 : 1) * C * LT * sizeof(real_t);
    static bool attr_set = false;
    if (!attr_set) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem); attr_set = true; }
    long long g = (nrows + R - 1) / R;
    const int cap = MINB * sms < kLgMaxGrid ? MINB * sms : kLgMaxGrid;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    kern<<<(unsigned) g, LT, smem, st>>>(X, ldx, y, sw, nrows, ncols, w, v, colpart, form);
    *nparts = (int) g;
}

template <int KIND>
bool launch_logistic_fused(const real_t* X, long long ldx, const real_t* y, const real_t* sw, long long nrows, long long ncols,
                           const real_t* w, const real_t* v, double* colpart, int* nparts, cudaStream_t st, LgForm form)
{
    const int cpt = (int) ((ncols + LT - 1) / LT);
    if (cpt > kLgMaxCpt) return false;
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms < 1) sms = 148; }
    static int variant = -1;
    if (variant < 0) { const char* e = getenv("STOCHQN_B200_LG_VARIANT"); variant = e ? atoi(e) : 0; }
#define LG_ARGS X, ldx, y, sw, nrows, ncols, w, v, colpart, nparts, sms, st, form
#define LG_CASE(C) if (cpt <= C) { launch_logistic_fused_t<C, KIND, lg_rows(C), 2>(LG_ARGS); return true; }
    // wide rows (measured on B200, tools/probe_logistic.py, STOCHQN_B200_LG_VARIANT is a dev switch):
    //   fp64, big batches : two CTAs per SM with 2 rows in flight each (6.0 TB/s at 50000 x 4097; a few spilled bytes)
    //   fp64, small batches: one CTA per SM with 4 rows in flight (fewer partial records to add up: 38 vs 49 us at 2000 x 4097)
    //   fp32              : two CTAs per SM, 4 rows in flight each (4-byte loads need more of them in flight)
#define LG_WIDE(C)                                                                                   \
    if (cpt <= C) {                                                                                  \
        if constexpr (sizeof(real_t) == 4) launch_logistic_fused_t<C, KIND, 4, 2>(LG_ARGS);          \
        else if (variant == 1 || (variant == 0 && nrows >= 8192)) launch_logistic_fused_t<C, KIND, 2, 2>(LG_ARGS); \
        else launch_logistic_fused_t<C, KIND, 4, 1>(LG_ARGS);                                        \
        return true;                                                                                 \
    }
    LG_CASE(1) LG_CASE(2) LG_CASE(4) LG_CASE(8) LG_CASE(12) LG_WIDE(16) LG_WIDE(17) LG_WIDE(20)
#undef LG_CASE
#undef LG_WIDE
#undef LG_ARGS
    return false;
}

int row_slices(long long nrows)
{
    long long s = nrows / 64;
    if (s < 1) s = 1;
    if (s > 64) s = 64;
    return (int) s;
}

template <bool SHARDED, int U, bool SHUFFLE>
static void launch_rosen_grad(const real_t* x, real_t* grad, long long n_local, long long offset, long long n_global,
                              const real_t* halo, const sqn::PeerArgs& pa, int* err, cudaStream_t st, real_t* halo_out = nullptr)
{
    constexpr int V = 16 / sizeof(real_t);
    if (((((uintptr_t) x) | ((uintptr_t) grad)) & 15u) == 0)
        rosen_grad_kernel<V, SHARDED, U, SHUFFLE><<<grid_1d(n_local / V), kT, 0, st>>>(x, grad, n_local, offset, n_global, halo, pa, err, halo_out);
    else
        rosen_grad_kernel<1, SHARDED, U, SHUFFLE><<<grid_1d(n_local), kT, 0, st>>>(x, grad, n_local, offset, n_global, halo, pa, err, halo_out);
}

template <bool SHARDED>
static void launch_rosen_grad_variant(const real_t* x, real_t* grad, long long n_local, long long offset, long long n_global,
                                      const real_t* halo, const sqn::PeerArgs& pa, int* err, cudaStream_t st, real_t* halo_out = nullptr)
{
    static int variant = -1;
    if (variant < 0) { const char* e = getenv("STOCHQN_B200_ROSEN_VARIANT"); variant = e ? atoi(e) : 0; }
    // measured on B200, n = 2^27 fp64 (STOCHQN_B200_ROSEN_VARIANT, dev switch): neighbours by L1-hitting scalar loads
    // with two chunks in flight 0.367 ms (5.85 TB/s); one chunk 0.445 ms; neighbours by warp shuffle 0.449-0.462 ms
    switch (variant) {
        case 1:  launch_rosen_grad<SHARDED, 1, true>(x, grad, n_local, offset, n_global, halo, pa, err, st, halo_out); break;
        case 2:  launch_rosen_grad<SHARDED, 2, true>(x, grad, n_local, offset, n_global, halo, pa, err, st, halo_out); break;
        case 3:  launch_rosen_grad<SHARDED, 1, false>(x, grad, n_local, offset, n_global, halo, pa, err, st, halo_out); break;
        case 4:  launch_rosen_grad<SHARDED, 4, false>(x, grad, n_local, offset, n_global, halo, pa, err, st, halo_out); break;
        default: launch_rosen_grad<SHARDED, 2, false>(x, grad, n_local, offset, n_global, halo, pa, err, st, halo_out); break;
    }
}

}  // namespace

extern "C" {

int stochqn_b200_rosenbrock_x0(real_t* x, long long n_local, long long offset, void* stream)
{
    rosen_x0_kernel<<<grid_1d(n_local), kT, 0, (cudaStream_t) stream>>>(x, n_local, offset);
    return check_launch("rosenbrock_x0");
}

int stochqn_b200_rosenbrock_grad(const real_t* x, real_t* grad, long long n_local, long long offset,
                                 long long n_global, const real_t* halo, void* stream)
{
    sqn::PeerArgs none;
    launch_rosen_grad_variant<false>(x, grad, n_local, offset, n_global, halo, none, nullptr, (cudaStream_t) stream);
    return check_launch("rosenbrock_grad");
}

int stochqn_b200_rosenbrock_grad_sharded(const real_t* x, real_t* grad, long long n_local, long long offset,
                                         long long n_global, int rank, int world_size, void* comm,
                                         real_t* halo, double* scratch, void* stream)
{
    if (world_size <= 1) return stochqn_b200_rosenbrock_grad(x, grad, n_local, offset, n_global, nullptr, stream);
    int* err = nullptr;
    sqn::PeerArgs pa = stochqn_b200_internal_next_exchange(comm, (size_t) 2 * world_size, &err, (cudaStream_t) stream);
    if (pa.world <= 1) {                                 // no peer-memory path: library all-reduce, then the plain kernel
        if (int r = stochqn_b200_rosenbrock_halo(x, n_local, rank, world_size, comm, halo, scratch, stream)) return r;
        return stochqn_b200_rosenbrock_grad(x, grad, n_local, offset, n_global, halo, stream);
    }
    // (the neighbours' edge values it fetched are left in `halo`: a re-evaluation at the same point can use the plain kernel)
    launch_rosen_grad_variant<true>(x, grad, n_local, offset, n_global, nullptr, pa, err, (cudaStream_t) stream, halo);
    return check_launch("rosenbrock_grad_sharded");
}

int stochqn_b200_rosenbrock_fun(const real_t* x, long long n_local, long long offset, long long n_global,
                                const real_t* halo, double* f_dev, void* stream)
{
    rosen_fun_kernel<<<grid_1d(n_local, 1024), kT, 0, (cudaStream_t) stream>>>(x, n_local, offset, n_global, halo, f_dev);
    return check_launch("rosenbrock_fun");
}

int stochqn_b200_rosenbrock_halo(const real_t* x, long long n_local, int rank, int world_size, void* comm,
                                 real_t* halo, double* scratch, void* stream)
{
    if (world_size > 512) return -1;
    rosen_halo_pack<<<1, 1024, 0, (cudaStream_t) stream>>>(x, n_local, rank, world_size, scratch);
    if (int r = stochqn_b200_allreduce_f64(comm, scratch, (size_t) 2 * world_size, stream)) return r;
    rosen_halo_unpack<<<1, 32, 0, (cudaStream_t) stream>>>(scratch, rank, world_size, halo);
    return check_launch("rosenbrock_halo", 2);
}

int stochqn_b200_csr_to_dense(const long long* indptr, const long long* indices, const real_t* data, long long row0, long long nrows,
                              long long ncols, real_t* out, long long ldo, int* bad_index_flag, void* stream)
{
    if (!indptr || !out || nrows < 0 || ncols <= 0 || ldo < ncols || row0 < 0) return -1;
    if (nrows == 0) return 0;
    csr_rows_to_dense_kernel<<<grid_1d(nrows * 32, 148 * 16), kT, 0, (cudaStream_t) stream>>>(indptr, indices, data, row0, nrows, ncols, out, ldo,
                                                                                              bad_index_flag);
    return check_launch("csr_to_dense");
}

size_t stochqn_b200_logistic_work_size(long long nrows, long long ncols)
{
    const long long two_sweep = (long long) row_slices(nrows) * ncols;
    const long long fused = (long long) kLgMaxGrid * (ncols + 2);
    return sizeof(double) * (size_t) (nrows + 2 + (two_sweep > fused ? two_sweep : fused));
}

static int logistic_common(int kind, const real_t* X, long long ldx, const real_t* y, const real_t* sw,
                           long long nrows, long long ncols, const real_t* w, const real_t* v, real_t lambda,
                           real_t* out, double* loss, void* work, cudaStream_t st, LgForm form = LgForm{0, 0})
{
    double* r = (double*) work;
    const long long nout = ncols + (form.icpt ? 1 : 0);
    double* colpart = r + nrows + 2;
    if (kind != LG_LOSS && !getenv("STOCHQN_B200_LOGISTIC_TWO_SWEEP")) {      // one sweep of the batch (ncols <= 5120)
        int nparts = 0;
        const bool done = kind == LG_GRAD ? launch_logistic_fused<LG_GRAD>(X, ldx, y, sw, nrows, ncols, w, v, colpart, &nparts, st, form)
                                          : launch_logistic_fused<LG_HVP>(X, ldx, y, sw, nrows, ncols, w, v, colpart, &nparts, st, form);
        if (done) {
            logistic_fused_finish<<<(unsigned) ((ncols + kFinCols - 1) / kFinCols), kFinCols * kFinSlices, 0, st>>>(colpart, nparts, ncols, kind == LG_HVP ? v : w, lambda, out, form);
            return check_launch("logistic (fused)", 2);
        }
    }
    const int g_rows = grid_1d(nrows * 32);
    if (kind == LG_GRAD) logistic_rows_kernel<LG_GRAD><<<g_rows, kT, 0, st>>>(X, ldx, y, sw, nrows, ncols, w, v, r, form);
    else if (kind == LG_HVP) logistic_rows_kernel<LG_HVP><<<g_rows, kT, 0, st>>>(X, ldx, y, sw, nrows, ncols, w, v, r, form);
    else logistic_rows_kernel<LG_LOSS><<<g_rows, kT, 0, st>>>(X, ldx, y, sw, nrows, ncols, w, v, r, form);
    logistic_norm_kernel<<<1, kT, 0, st>>>(sw, nrows, r);
    if (kind == LG_LOSS) {
        logistic_loss_finish_kernel<<<1, kT, 0, st>>>(r, nrows, w, ncols, lambda, loss, form);
    } else {
        const int slices = row_slices(nrows);
        dim3 grid((unsigned) ((ncols + kT - 1) / kT), (unsigned) slices);
        logistic_cols_kernel<<<grid, kT, 0, st>>>(X, ldx, nrows, ncols, r, colpart);
        logistic_finish_kernel<<<(unsigned) ((nout + kT - 1) / kT), kT, 0, st>>>(colpart, slices, nrows, ncols, r,
                                                                                 kind == LG_HVP ? v : w, lambda, out, form);
    }
    return check_launch("logistic", kind == LG_LOSS ? 3 : 4);
}

int stochqn_b200_logistic_grad(const real_t* X, long long ldx, const real_t* y, const real_t* sw, long long nrows,
                               long long ncols, const real_t* w, real_t lambda, real_t* grad, void* work, void* stream)
{
    return logistic_common(LG_GRAD, X, ldx, y, sw, nrows, ncols, w, nullptr, lambda, grad, nullptr, work, (cudaStream_t) stream);
}

int stochqn_b200_logistic_hess_vec(const real_t* X, long long ldx, const real_t* y, const real_t* sw, long long nrows,
                                   long long ncols, const real_t* w, const real_t* v, real_t lambda, real_t* hess_vec,
                                   void* work, void* stream)
{
    return logistic_common(LG_HVP, X, ldx, y, sw, nrows, ncols, w, v, lambda, hess_vec, nullptr, work, (cudaStream_t) stream);
}

int stochqn_b200_logistic_loss(const real_t* X, long long ldx, const real_t* y, const real_t* sw, long long nrows,
                               long long ncols, const real_t* w, real_t lambda, double* loss_dev, void* work, void* stream)
{
    return logistic_common(LG_LOSS, X, ldx, y, sw, nrows, ncols, w, nullptr, lambda, nullptr, loss_dev, work, (cudaStream_t) stream);
}

// scikit-learn (<= 1.0) conventions: see LgForm above and include/stochqn_b200.h
int stochqn_b200_logistic_sk_grad(const real_t* X, long long ldx, const real_t* y, const real_t* sw, long long nrows,
                                  long long ncols, int fit_intercept, const real_t* w, real_t alpha, real_t* grad,
                                  void* work, void* stream)
{
    return logistic_common(LG_GRAD, X, ldx, y, sw, nrows, ncols, w, nullptr, alpha, grad, nullptr, work, (cudaStream_t) stream,
                           LgForm{1, fit_intercept ? 1 : 0});
}

int stochqn_b200_logistic_sk_hess_vec(const real_t* X, long long ldx, const real_t* y, const real_t* sw, long long nrows,
                                      long long ncols, int fit_intercept, const real_t* w, const real_t* v, real_t alpha,
                                      real_t* hess_vec, void* work, void* stream)
{
    return logistic_common(LG_HVP, X, ldx, y, sw, nrows, ncols, w, v, alpha, hess_vec, nullptr, work, (cudaStream_t) stream,
                           LgForm{1, fit_intercept ? 1 : 0});
}

int stochqn_b200_logistic_sk_loss(const real_t* X, long long ldx, const real_t* y, const real_t* sw, long long nrows,
                                  long long ncols, int fit_intercept, const real_t* w, real_t alpha, double* loss_dev,
                                  void* work, void* stream)
{
    return logistic_common(LG_LOSS, X, ldx, y, sw, nrows, ncols, w, nullptr, alpha, nullptr, loss_dev, work, (cudaStream_t) stream,
                           LgForm{1, fit_intercept ? 1 : 0});
}

}  // extern "C"
