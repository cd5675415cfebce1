// multinomial.cu - bundled device callbacks for multinomial (softmax) logistic regression: loss, gradient and
// Hessian-vector product with the semantics of the scikit-learn (<= 1.0) private functions the reference's Python
// layer calls (stochqn/_logistic.py:7-13 -> _multinomial_loss_grad / _multinomial_grad_hess):
//
//   W = w reshaped (K, d [+1]) row-major, intercept = last column;   Z = X W' + b ;  P = softmax(Z) (rows)
//   loss      = -sum_i sw_i sum_k Y_ik log P_ik + alpha/2 ||W||^2                     (sums, not means)
//   grad      = (sw .* (P - Y))' X + alpha W ;  intercept column = column sums of sw .* (P - Y)
//   hess_vec  : R = X V' + vb ;  R += rowsum(-P .* R) ;  R .*= P ;  R .*= sw ;  out = R' X + alpha V ; intercept: column sums of R
//
// Data flow (all on the caller's stream, no host round trip):
//   mn_transpose   XT = X_batch'                                   (so that BOTH products are "TN" GEMMs: C = A B', K contiguous)
//   GEMM 1         Z  = X_batch  W(:, :d)'      [B x K]            (and R = X_batch V(:, :d)' for the Hessian-vector product)
//   mn_rows        one warp per sample: log-sum-exp, loss term, D = sw .* (P - Y)  (or the R-operator row) -> DT = D' [K x B]
//   GEMM 2         G  = DT XT'                  [K x d]
//   mn_finish      G += alpha W ;  intercept column = row sums of DT ;  loss = sum of the per-sample terms + alpha/2 ||W||^2
// The GEMMs are the only compute-bound work of the whole path (2 B d K flop each).  gemm_tn below is the portable
// CUDA-core version (any shape, fp64 and fp32, split-K for the skinny config-3 shapes); the fp32 build routes large
// aligned products to the tcgen05 / TMEM tensor-core kernel of gemm_tf32_sm100.cuh.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <atomic>

#include "stochqn.h"
#include "stochqn_b200.h"

extern std::atomic<unsigned long long> stochqn_b200_cb_launches;    // callbacks.cu

#ifdef USE_FLOAT
#include "gemm_tf32_sm100.cuh"
#endif

#include "internal_rs.h"

namespace {

// where the optional fused reduce-scatter of the gradient product sends its elements (see gemm_tf32_sm100.cuh)
struct MnScatter {
    int world = 0;
    long long blk = 0;
    void* dst[16] = {};
};

constexpr int TM = 64, TN = 64, TK = 16, NT = 256;

int mn_check(const char* what, int launched)
{
    stochqn_b200_cb_launches.fetch_add((unsigned long long) launched, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        fprintf(stderr, "stochqn_b200: %s launch failed: %s\n", what, cudaGetErrorString(e));
        return -2;
    }
    return 0;
}

// C_z[M x N] (ldc) = A[M x Kc](lda) * B[N x Kc](ldb)' restricted to the k-range of blockIdx.z; C_z = C + z * cz_stride.
// 64 x 64 tile, 16-deep slices, 256 threads x (4 x 4) accumulators in the storage type.
template <typename T>
__global__ void __launch_bounds__(NT)
gemm_tn(const T* __restrict__ A, long long lda, const T* __restrict__ B, long long ldb, T* __restrict__ C, long long ldc,
        long long cz_stride, int M, int N, int Kc, int kper)
{
    __shared__ T As[TK][TM + 4];
    __shared__ T Bs[TK][TN + 4];
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    const int k0 = blockIdx.z * kper;
    const int k1 = (k0 + kper < Kc) ? k0 + kper : Kc;
    const int lrow = threadIdx.x / 4, lk = (threadIdx.x % 4) * 4;
    T acc[4][4];
    #pragma unroll
    for (int i = 0; i < 4; ++i)
        #pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = (T) 0;
    for (int kk = k0; kk < k1; kk += TK) {
        #pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = kk + lk + i;
            As[lk + i][lrow] = (m0 + lrow < M && k < k1) ? A[(long long) (m0 + lrow) * lda + k] : (T) 0;
            Bs[lk + i][lrow] = (n0 + lrow < N && k < k1) ? B[(long long) (n0 + lrow) * ldb + k] : (T) 0;
        }
        __syncthreads();
        #pragma unroll
        for (int k = 0; k < TK; ++k) {
            T a[4], b[4];
            #pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[k][ty * 4 + i]; b[i] = Bs[k][tx * 4 + i]; }
            #pragma unroll
            for (int i = 0; i < 4; ++i)
                #pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    T* Cz = C + (long long) blockIdx.z * cz_stride;
    #pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
        #pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < N) Cz[(long long) m * ldc + n] = acc[i][j];
        }
    }
}

// XT[d x ldt] = X[B x ldx]'  (32 x 32 tiles through shared memory, both sides coalesced)
template <typename T>
__global__ void mn_transpose(const T* __restrict__ X, long long ldx, T* __restrict__ XT, long long ldt, int B, int d)
{
    __shared__ T tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;      // bx: feature block, by: sample block
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int i = by + r, j = bx + threadIdx.x;
        tile[r][threadIdx.x] = (i < B && j < d) ? X[(long long) i * ldx + j] : (T) 0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int j = bx + r, i = by + threadIdx.x;
        if (j < d && i < B) XT[(long long) j * ldt + i] = tile[threadIdx.x][r];
    }
}

enum { MN_GRAD = 0, MN_HVP = 1 };

// One warp per sample, 8 samples per CTA.  Z (and R) arrive as `splits` partial products that are added here in a
// fixed order together with the intercepts.  Writes DT[k][i] (transposed through shared memory so that the stores are
// contiguous over samples) and the per-sample loss term.
template <typename T, int KIND>
__global__ void __launch_bounds__(256)
mn_rows(const T* __restrict__ Zp, const T* __restrict__ Rp, int splits, long long zstride, int B, int K,
        const T* __restrict__ wb, long long ldw, const T* __restrict__ vb, const T* __restrict__ Y, long long ldy,
        const int* __restrict__ labels, const T* __restrict__ sw, T* __restrict__ DT, long long ldt,
        double* __restrict__ loss_terms)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i0 = blockIdx.x * 8, i = i0 + warp;
    const bool live = i < B;
    __shared__ T tile[8][33];
    auto zval = [&](int k) -> double {
        double z = wb ? (double) wb[(long long) k * ldw] : 0.0;
        for (int s = 0; s < splits; ++s) z += (double) Zp[(long long) s * zstride + (long long) i * K + k];
        return z;
    };
    auto rval = [&](int k) -> double {
        double r = vb ? (double) vb[(long long) k * ldw] : 0.0;
        for (int s = 0; s < splits; ++s) r += (double) Rp[(long long) s * zstride + (long long) i * K + k];
        return r;
    };
    auto warp_max = [&](double v) { for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o)); return v; };
    auto warp_add = [&](double v) { for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); return v; };
    double lse = 0.0, wt = 1.0, pr = 0.0;
    if (live) {
        double mx = -INFINITY;
        for (int k = lane; k < K; k += 32) mx = fmax(mx, zval(k));
        mx = warp_max(mx);
        double se = 0.0;
        for (int k = lane; k < K; k += 32) se += exp(zval(k) - mx);
        se = warp_add(se);
        lse = mx + log(se);
        wt = sw ? (double) sw[i] : 1.0;
        if (loss_terms) {                                      // -sw_i sum_k Y_ik (z_ik - lse)
            double lt = 0.0;
            if (labels) { if (lane == 0) lt = -(zval(labels[i]) - lse); }
            else for (int k = lane; k < K; k += 32) { const double yk = (double) Y[(long long) i * ldy + k]; if (yk != 0.0) lt -= yk * (zval(k) - lse); }
            lt = warp_add(lt);
            if (lane == 0) loss_terms[i] = wt * lt;
        }
        if (KIND == MN_HVP) {
            for (int k = lane; k < K; k += 32) pr += exp(zval(k) - lse) * rval(k);
            pr = warp_add(pr);
        }
    }
    if (!DT) return;
    for (int k0 = 0; k0 < K; k0 += 32) {                       // CTA-uniform trip count
        const int k = k0 + lane;
        T out = (T) 0;
        if (live && k < K) {
            const double p = exp(zval(k) - lse);
            if (KIND == MN_GRAD) {
                const double yk = labels ? (labels[i] == k ? 1.0 : 0.0) : (double) Y[(long long) i * ldy + k];
                out = (T) (wt * (p - yk));
            } else {
                out = (T) (wt * p * (rval(k) - pr));
            }
        }
        tile[warp][lane] = out;
        __syncthreads();
        const int kk = threadIdx.x / 8, r = threadIdx.x % 8;   // 8 consecutive threads -> 8 consecutive samples
        if (k0 + kk < K && i0 + r < B) DT[(long long) (k0 + kk) * ldt + i0 + r] = tile[r][kk];
        __syncthreads();
    }
}

// ---- the same row work for large batches x many classes, as two well-occupied kernels -----------------------------
// mn_rows gives one warp a whole sample: at B = 1024, K = 4096 that is 1024 warps on the whole GPU, three strided passes
// over the intercept column of W per sample, and 32-byte store segments - 297 us between two 115 us GEMMs.  Here:
//   mn_gather_col  the intercept columns of W (and V) gathered once into contiguous vectors
//   mn_stats       one CTA per sample: log-sum-exp (and sum_k p_k r_k for the Hessian-vector product), loss term
//   mn_dtile       32 x 32 tiles: read Z along classes, write DT along samples (both coalesced)
template <typename T>
__global__ void __launch_bounds__(256)
mn_gather_col(const T* __restrict__ W, const T* __restrict__ V, long long ldw, int K, T* __restrict__ bw, T* __restrict__ bv)
{
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (k >= K) return;
    bw[k] = W[(long long) k * ldw];
    if (V) bv[k] = V[(long long) k * ldw];
}

__device__ __forceinline__ double mn_block_sum(double v, double* red)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0;
    #pragma unroll
    for (int q = 0; q < 8; ++q) t += red[q];
    __syncthreads();
    return t;
}
__device__ __forceinline__ double mn_block_max(double v, double* red)
{
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = red[0];
    #pragma unroll
    for (int q = 1; q < 8; ++q) t = fmax(t, red[q]);
    __syncthreads();
    return t;
}

template <typename T, int KIND>
__global__ void __launch_bounds__(256)
mn_stats(const T* __restrict__ Zp, const T* __restrict__ Rp, int splits, long long zstride, int B, int K,
         const T* __restrict__ bw, const T* __restrict__ bv, const T* __restrict__ Y, long long ldy,
         const int* __restrict__ labels, const T* __restrict__ sw, double* __restrict__ stats, double* __restrict__ loss_terms)
{
    __shared__ double red[8];
    const int i = blockIdx.x;
    auto zval = [&](int k) -> double {
        double z = bw ? (double) bw[k] : 0.0;
        for (int s = 0; s < splits; ++s) z += (double) Zp[(long long) s * zstride + (long long) i * K + k];
        return z;
    };
    double mx = -INFINITY;
    for (int k = threadIdx.x; k < K; k += 256) mx = fmax(mx, zval(k));
    mx = mn_block_max(mx, red);
    double se = 0.0;
    for (int k = threadIdx.x; k < K; k += 256) se += exp(zval(k) - mx);
    se = mn_block_sum(se, red);
    const double lse = mx + log(se);
    if (loss_terms) {                                          // -sw_i sum_k Y_ik (z_ik - lse)
        const double wt = sw ? (double) sw[i] : 1.0;
        double lt = 0.0;
        if (labels) { if (threadIdx.x == 0) lt = -(zval(labels[i]) - lse); }
        else for (int k = threadIdx.x; k < K; k += 256) { const double yk = (double) Y[(long long) i * ldy + k]; if (yk != 0.0) lt -= yk * (zval(k) - lse); }
        lt = mn_block_sum(lt, red);
        if (threadIdx.x == 0) loss_terms[i] = wt * lt;
    }
    double pr = 0.0;
    if (KIND == MN_HVP) {
        for (int k = threadIdx.x; k < K; k += 256) {
            double r = bv ? (double) bv[k] : 0.0;
            for (int s = 0; s < splits; ++s) r += (double) Rp[(long long) s * zstride + (long long) i * K + k];
            pr += exp(zval(k) - lse) * r;
        }
        pr = mn_block_sum(pr, red);
    }
    if (threadIdx.x == 0) { stats[2 * (long long) i] = lse; stats[2 * (long long) i + 1] = pr; }
}

template <typename T, int KIND>
__global__ void __launch_bounds__(256)
mn_dtile(const T* __restrict__ Zp, const T* __restrict__ Rp, int splits, long long zstride, int B, int K,
         const T* __restrict__ bw, const T* __restrict__ bv, const T* __restrict__ Y, long long ldy,
         const int* __restrict__ labels, const T* __restrict__ sw, const double* __restrict__ stats,
         T* __restrict__ DT, long long ldt)
{
    __shared__ T tile[32][33];
    const int k0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int i = i0 + r, k = k0 + tx;
        T out = (T) 0;
        if (i < B && k < K) {
            double z = bw ? (double) bw[k] : 0.0;
            for (int s = 0; s < splits; ++s) z += (double) Zp[(long long) s * zstride + (long long) i * K + k];
            const double wt = sw ? (double) sw[i] : 1.0;
            const double p = exp(z - stats[2 * (long long) i]);
            if (KIND == MN_GRAD) {
                const double yk = labels ? (labels[i] == k ? 1.0 : 0.0) : (double) Y[(long long) i * ldy + k];
                out = (T) (wt * (p - yk));
            } else {
                double rr = bv ? (double) bv[k] : 0.0;
                for (int s = 0; s < splits; ++s) rr += (double) Rp[(long long) s * zstride + (long long) i * K + k];
                out = (T) (wt * p * (rr - stats[2 * (long long) i + 1]));
            }
        }
        tile[r][tx] = out;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int k = k0 + r, i = i0 + tx;
        if (k < K && i < B) DT[(long long) k * ldt + i] = tile[tx][r];
    }
}

// One CTA per class k: G[k][0..d) += alpha * U[k][0..d) ;  G[k][d] = sum_i DT[k][i] (when there is an intercept).
template <typename T>
__global__ void __launch_bounds__(256)
mn_finish(T* __restrict__ G, long long ldg, const T* __restrict__ U, long long ldu, int d, int fit_intercept, T alpha,
          const T* __restrict__ DT, long long ldt, int B)
{
    const int k = blockIdx.x;
    for (int j = threadIdx.x; j < d; j += blockDim.x) G[(long long) k * ldg + j] = fma(alpha, U[(long long) k * ldu + j], G[(long long) k * ldg + j]);
    if (!fit_intercept) return;
    double s = 0.0;
    for (int i = threadIdx.x; i < B; i += blockDim.x) s += (double) DT[(long long) k * ldt + i];
    __shared__ double red[8];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int q = 0; q < 8; ++q) t += red[q];
        G[(long long) k * ldg + d] = (T) t;
    }
}

// wn[k] = sum_{j<d} W[k][j]^2 : one CTA per class (the penalty term of the loss, without the intercept column)
template <typename T>
__global__ void __launch_bounds__(256)
mn_wnorm(const T* __restrict__ W, long long ldw, int d, double* __restrict__ wn)
{
    const T* row = W + (long long) blockIdx.x * ldw;
    double b = 0.0;
    for (int j = threadIdx.x; j < d; j += 256) { const double w = (double) row[j]; b = fma(w, w, b); }
    __shared__ double rb[8];
    for (int o = 16; o > 0; o >>= 1) b += __shfl_down_sync(0xffffffffu, b, o);
    if ((threadIdx.x & 31) == 0) rb[threadIdx.x >> 5] = b;
    __syncthreads();
    if (threadIdx.x == 0) {
        double sb = 0;
        for (int q = 0; q < 8; ++q) sb += rb[q];
        wn[blockIdx.x] = sb;
    }
}

// loss = sum_i terms[i] + alpha/2 * sum_k wn[k] ; one CTA, fixed order
__global__ void __launch_bounds__(256)
mn_loss_finish(const double* __restrict__ terms, int B, const double* __restrict__ wn, int K, double alpha, double* __restrict__ loss)
{
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < B; i += 256) a += terms[i];
    for (int k = threadIdx.x; k < K; k += 256) b += wn[k];
    __shared__ double ra[8], rb[8];
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_down_sync(0xffffffffu, a, o); b += __shfl_down_sync(0xffffffffu, b, o); }
    if ((threadIdx.x & 31) == 0) { ra[threadIdx.x >> 5] = a; rb[threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double sa = 0, sb = 0;
        for (int q = 0; q < 8; ++q) { sa += ra[q]; sb += rb[q]; }
        *loss = sa + 0.5 * alpha * sb;
    }
}

struct MnPlan {
    int splits;
    long long bpad;                   // leading dimension of DT / XT (samples, padded to 16 bytes)
    long long dpad;                   // leading dimension of the packed coefficient copies (features, padded to 16 bytes)
    size_t off_zp, off_rp, off_dt, off_xt, off_terms, off_stats, off_bvec, off_wp, off_vp, off_small, off_small_bar, total;
    bool small_ok;                    // the work buffer holds the partial products of the one-launch small-batch gradient
};

// Wp[K x ldp] = W[:, :d]  (coefficient block without the intercept column, rows 16-byte aligned: what TMA needs)
// One CTA per coefficient row at a time (no index division; coalesced 4-byte accesses on both sides).
template <typename T>
__global__ void __launch_bounds__(256)
mn_pack(const T* __restrict__ W, long long ldw, T* __restrict__ Wp, long long ldp, int K, int d)
{
    for (int k = blockIdx.x; k < K; k += gridDim.x) {
        const T* src = W + (long long) k * ldw;
        T* dst = Wp + (long long) k * ldp;
        for (int j = threadIdx.x; j < d; j += 256) dst[j] = __ldg(src + j);
    }
}

// intercept column of the gradient / Hessian-vector product: G[k][d] = sum_i DT[k][i]; one warp per class
template <typename T>
__global__ void __launch_bounds__(256)
mn_intercept(T* __restrict__ G, long long ldg, int d, const T* __restrict__ DT, long long ldt, int B, int K, const MnScatter sc)
{
    const int lane = threadIdx.x & 31;
    const int k = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (k >= K) return;
    double s = 0.0;
    for (int i = lane; i < B; i += 32) s += (double) DT[(long long) k * ldt + i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane != 0) return;
    const long long e = (long long) k * ldg + d;
    if (sc.world > 1) {                                    // fused reduce-scatter: straight into the owner's receive slot
        const int o = (int) (e / sc.blk);
        if (o < sc.world) reinterpret_cast<T*>(sc.dst[o])[e - (long long) o * sc.blk] = (T) s;
    } else G[e] = (T) s;
}

// receive side of the fused reduce-scatter: out[e] = sum over senders (rank order) of slot[s][e]
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
mn_rs_reduce(const T* __restrict__ slots, long long blk, int world, T* __restrict__ out)
{
    const long long nv = blk / VEC;
    const long long stride = (long long) gridDim.x * blockDim.x;
    for (long long v = (long long) blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += stride) {
        if constexpr (VEC == 4) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int r = 0; r < world; ++r) {
                const float4 t = __ldcg(reinterpret_cast<const float4*>(slots + (long long) r * blk) + v);
                acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
            }
            reinterpret_cast<float4*>(out)[v] = acc;
        } else {
            T acc = (T) 0;
            for (int r = 0; r < world; ++r) acc += __ldcg(slots + (long long) r * blk + v);
            out[v] = acc;
        }
    }
}

MnPlan mn_plan(long long B, long long d, long long K);

// ---- small batches: the whole gradient in ONE cooperative launch (mn_small.cuh) ----
}  // namespace
#include "mn_small.cuh"
namespace {
using namespace mnsmall;

unsigned long long* g_mn_trace = nullptr;

size_t mn_small_smem(long long B, long long d, long long K, int grid)
{
    const long long cw = (d + grid - 1) / grid;
    size_t ds = sizeof(real_t) * (size_t) (B * K);
    if (ds < 2048 * sizeof(double)) ds = 2048 * sizeof(double);        // phase 2 stages its partial sums there ([slices][padded row])
    return sizeof(real_t) * (size_t) (cw * B + cw * K) + ds + 16;
}

// 0: launched; 1: not applicable (the caller takes the multi-launch route); < 0: error
int mn_try_small(const real_t* X, long long ldx, const real_t* Y, long long ldy, const int* labels, const real_t* sw,
                 long long B, long long d, long long K, int fit_intercept, const real_t* w, real_t alpha, real_t* out,
                 unsigned char* base, const MnPlan& p, cudaStream_t st)
{
    static int enabled = -1, sms = 0, coop = 0;
    if (enabled < 0) {
        const char* e = getenv("STOCHQN_B200_MN_SMALL");
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        enabled = (e && atoi(e) == 0) ? 0 : 1;
    }
    if (!enabled || !coop || sms < 1 || sms > MS_MAX_GRID || !p.small_ok) return 1;
    if (B > MS_MAX_B || K > MS_MAX_K || B * K > MS_MAX_BK || (d + sms - 1) / sms > MS_CW) return 1;
    const size_t smem = mn_small_smem(B, d, K, sms);
    if (smem > 200 * 1024) return 1;
    auto kern = mn_grad_small<real_t>;
    static size_t smem_set = 0;
    if (smem > smem_set) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (200 * 1024)) != cudaSuccess) { cudaGetLastError(); return 1; }
        smem_set = 200 * 1024;
    }
    real_t* Zp = (real_t*) (base + p.off_small);
    real_t* Dg = (real_t*) (base + p.off_dt);
    unsigned long long* bar = (unsigned long long*) (base + p.off_small_bar);
    if (cudaMemsetAsync(bar, 0, 64 * sizeof(unsigned long long), st) != cudaSuccess) return -2;
    MnSmallArgs<real_t> ma;
    ma.X = X; ma.ldx = ldx; ma.Y = Y; ma.ldy = ldy; ma.labels = labels; ma.sw = sw;
    ma.B = (int) B; ma.d = (int) d; ma.K = (int) K; ma.icpt = fit_intercept ? 1 : 0;
    ma.W = w; ma.alpha = alpha; ma.Gout = out; ma.Zp = Zp; ma.Dg = Dg;
    unsigned long long* trace = g_mn_trace;
    void* args[] = {&ma, &bar, &trace};
    if (cudaLaunchCooperativeKernel((const void*) kern, dim3((unsigned) sms), dim3(MS_T), args, smem, st) != cudaSuccess) {
        fprintf(stderr, "stochqn_b200: mn_grad_small launch failed: %s\n", cudaGetErrorString(cudaGetLastError()));
        return -2;
    }
    return mn_check("mn_grad_small", 1);
}

MnPlan mn_plan(long long B, long long d, long long K)
{
    MnPlan p;
    const long long tiles = ((B + TM - 1) / TM) * ((K + TN - 1) / TN);
    long long s = tiles >= 148 ? 1 : (296 + tiles - 1) / tiles;
    const long long smax = (d + 4 * TK - 1) / (4 * TK);
    if (s > smax) s = smax;
    if (s > 32) s = 32;
    if (s < 1) s = 1;
    p.splits = (int) s;
    const long long q = 16 / (long long) sizeof(real_t);
    p.bpad = (B + q - 1) / q * q;
    p.dpad = (d + q - 1) / q * q;
    auto al = [](size_t v) { return (v + 255) / 256 * 256; };
    size_t off = 0;
    // The split-k partial products take s * B * K elements, and s is not monotone in B (fewer splits once the tiles fill the GPU):
    // a buffer sized for a long batch must also hold the plan of every shorter one, so the regions are sized by a bound that IS
    // monotone in B:  s * tiles <= 296 + tiles  =>  s * B * K <= 296 * 64 * 64 + (B + 64) * (K + 64).
    const size_t zp_elems = (size_t) 296 * TM * TN + (size_t) (B + TM) * (size_t) (K + TN);
    p.off_zp = off; off = al(off + sizeof(real_t) * zp_elems);
    p.off_rp = off; off = al(off + sizeof(real_t) * zp_elems);
    p.off_dt = off; off = al(off + sizeof(real_t) * (size_t) (K * p.bpad));
    p.off_xt = off; off = al(off + sizeof(real_t) * (size_t) (d * p.bpad));
    p.off_terms = off; off = al(off + sizeof(double) * (size_t) (B + K + 2));       // per-sample loss terms, then per-class ||w_k||^2
    p.off_stats = off; off = al(off + sizeof(double) * (size_t) (2 * B));           // per-sample log-sum-exp and sum_k p_k r_k
    p.off_bvec = off; off = al(off + sizeof(real_t) * (size_t) (2 * K));            // intercept columns of W and V, gathered
    p.off_wp = off; p.off_vp = off;
#ifdef USE_FLOAT
    p.off_wp = off; off = al(off + sizeof(real_t) * (size_t) (K * p.dpad));      // packed W / V for the tensor-core path
    p.off_vp = off; off = al(off + sizeof(real_t) * (size_t) (K * p.dpad));
#endif
    // one-launch small-batch gradient (mn_grad_small): one partial Z per CTA + the barrier words.  Reserved for every
    // batch size (capped at its largest applicable one), so a buffer sized for a long batch also serves the mini-batches.
    p.small_ok = K <= MS_MAX_K;
    p.off_small = off; p.off_small_bar = off;
    if (p.small_ok) {
        long long bk = (B < MS_MAX_B ? B : MS_MAX_B) * K;
        if (bk > MS_MAX_BK) bk = MS_MAX_BK;
        const size_t dt_need = sizeof(real_t) * (size_t) bk;                 // D is kept where DT lives: K * bpad >= B * K
        (void) dt_need;
        off = al(off + sizeof(real_t) * (size_t) (MS_MAX_GRID * (bk + 3 * MS_MAX_B)));      // rows padded to 16 bytes
        p.off_small_bar = off; off = al(off + 64 * sizeof(unsigned long long));
    }
    p.total = off;
    return p;
}

// C = A B' with the k-range split `splits` ways (partials `cz_stride` apart); tensor cores when the fp32 build can
// `addend` / `alpha`: C = A B' + alpha * addend fused into the tensor-core epilogue; *fused tells the caller whether that
// happened (the CUDA-core kernel leaves it to a separate pass).
int launch_gemm(const real_t* A, long long lda, const real_t* Bm, long long ldb, real_t* C, long long ldc, long long cz_stride,
                int M, int N, int Kc, int splits, cudaStream_t st, const real_t* addend = nullptr, long long ld_add = 0,
                real_t alpha = 0, bool* fused = nullptr, const MnScatter* scatter = nullptr)
{
    if (fused) *fused = false;
#ifdef USE_FLOAT
    if (splits == 1 && !getenv("STOCHQN_B200_NO_TENSOR_CORES") && sm100_gemm_tf32_usable(A, lda, Bm, ldb, C, ldc, M, N, Kc)) {
        tf32gemm::ScatterArgs sa;
        if (scatter && scatter->world > 1) {
            sa.world = scatter->world; sa.blk = scatter->blk;
            for (int r = 0; r < scatter->world && r < 16; ++r) sa.dst[r] = (float*) scatter->dst[r];
        }
        const int r = sm100_gemm_tf32(A, lda, Bm, ldb, C, ldc, M, N, Kc, st, addend, ld_add, alpha, sa.world > 1 ? &sa : nullptr);
        if (r == 0) { if (fused) *fused = addend != nullptr; return mn_check("gemm (tcgen05 tf32)", 1); }
        if (r != 1) return r;                       // 1 = not available on this device / driver: fall through
    }
#endif
    if (scatter && scatter->world > 1) return -5;   // the fused reduce-scatter exists in the tensor-core epilogue only
    const int kper = ((Kc + splits - 1) / splits + TK - 1) / TK * TK;
    dim3 grid((unsigned) ((N + TN - 1) / TN), (unsigned) ((M + TM - 1) / TM), (unsigned) splits);
    gemm_tn<real_t><<<grid, NT, 0, st>>>(A, lda, Bm, ldb, C, ldc, cz_stride, M, N, Kc, kper);
    return mn_check("gemm_tn", 1);
}

int mn_common(int kind, const real_t* X, long long ldx, const real_t* Y, long long ldy, const int* labels, const real_t* sw,
              long long B, long long d, long long K, int fit_intercept, const real_t* w, const real_t* v, real_t alpha,
              real_t* out, double* loss_dev, void* work, cudaStream_t st, const MnScatter* scatter = nullptr)
{
    if (B <= 0 || d <= 0 || K <= 0 || (kind == MN_GRAD && !Y && !labels) || !X || !w || !work) return -1;
    if (B > 2000000000ll || d > 2000000000ll || K > 2000000000ll) return -1;
    const MnPlan p = mn_plan(B, d, K);
    unsigned char* base = (unsigned char*) work;
    real_t* Zp = (real_t*) (base + p.off_zp);
    real_t* Rp = (real_t*) (base + p.off_rp);
    real_t* DT = (real_t*) (base + p.off_dt);
    real_t* XT = (real_t*) (base + p.off_xt);
    double* terms = (double*) (base + p.off_terms);
    const long long ldw = d + (fit_intercept ? 1 : 0);
    const long long zstride = B * K;
    const bool need_out = out != nullptr;
    const real_t *w1 = w, *v1 = v;
    long long ld1 = ldw;
#ifdef USE_FLOAT
    // the coefficient rows are (d + intercept) floats apart - not 16-byte aligned in general: the tensor-core path
    // reads a packed copy (one extra pass over W, small against 2*B*d*K flop)
    if (p.splits == 1 && (ldw & 3) && !getenv("STOCHQN_B200_NO_TENSOR_CORES") &&
        sm100_gemm_tf32_usable(X, ldx, (const real_t*) (base + p.off_wp), p.dpad, nullptr, 0, (int) B, (int) K, (int) d)) {
        real_t* Wp = (real_t*) (base + p.off_wp);
        mn_pack<real_t><<<(unsigned) (K < 4736 ? K : 4736), 256, 0, st>>>(w, ldw, Wp, p.dpad, (int) K, (int) d);
        w1 = Wp; ld1 = p.dpad;
        int packed = 1;
        if (kind == MN_HVP) {
            real_t* Vp = (real_t*) (base + p.off_vp);
            mn_pack<real_t><<<(unsigned) (K < 4736 ? K : 4736), 256, 0, st>>>(v, ldw, Vp, p.dpad, (int) K, (int) d);
            v1 = Vp;
            ++packed;
        }
        if (int r = mn_check("multinomial pack", packed)) return r;
    }
#endif
    if (kind == MN_GRAD && need_out && !loss_dev && !scatter) {          // small batches: the whole gradient in one launch
        const int r = mn_try_small(X, ldx, Y, ldy, labels, sw, B, d, K, fit_intercept, w, alpha, out, base, p, st);
        if (r <= 0) return r;
    }
    // GEMM 1: Z = X W'   (and R = X V')
    if (int r = launch_gemm(X, ldx, w1, ld1, Zp, K, zstride, (int) B, (int) K, (int) d, p.splits, st)) return r;
    if (kind == MN_HVP) { if (int r = launch_gemm(X, ldx, v1, ld1, Rp, K, zstride, (int) B, (int) K, (int) d, p.splits, st)) return r; }
    const real_t* wb = fit_intercept ? w + d : nullptr;
    const real_t* vb = (fit_intercept && v) ? v + d : nullptr;
    int launched = 1;
    if (B * K >= (1ll << 18) && B <= 65535ll * 32) {
        // large batch x many classes: gather the intercepts, per-sample statistics with one CTA per sample, tiled write of DT
        double* stats = (double*) (base + p.off_stats);
        real_t* bw = (real_t*) (base + p.off_bvec);
        real_t* bv = bw + K;
        if (wb) { mn_gather_col<real_t><<<(unsigned) ((K + 255) / 256), 256, 0, st>>>(wb, vb, ldw, (int) K, bw, bv); ++launched; }
        const real_t* bwp = wb ? bw : nullptr;
        const real_t* bvp = vb ? bv : nullptr;
        double* lt = (loss_dev && kind == MN_GRAD) ? terms : nullptr;
        const bool want_dt = kind == MN_HVP || need_out;
        dim3 tg((unsigned) ((K + 31) / 32), (unsigned) ((B + 31) / 32));
        if (kind == MN_GRAD) {
            mn_stats<real_t, MN_GRAD><<<(unsigned) B, 256, 0, st>>>(Zp, Rp, p.splits, zstride, (int) B, (int) K, bwp, bvp, Y, ldy, labels, sw, stats, lt);
            if (want_dt) { mn_dtile<real_t, MN_GRAD><<<tg, 256, 0, st>>>(Zp, Rp, p.splits, zstride, (int) B, (int) K, bwp, bvp, Y, ldy, labels, sw, stats, DT, p.bpad); ++launched; }
        } else {
            mn_stats<real_t, MN_HVP><<<(unsigned) B, 256, 0, st>>>(Zp, Rp, p.splits, zstride, (int) B, (int) K, bwp, bvp, Y, ldy, labels, sw, stats, nullptr);
            mn_dtile<real_t, MN_HVP><<<tg, 256, 0, st>>>(Zp, Rp, p.splits, zstride, (int) B, (int) K, bwp, bvp, Y, ldy, labels, sw, stats, DT, p.bpad);
            ++launched;
        }
    } else {
        const unsigned rows_grid = (unsigned) ((B + 7) / 8);
        if (kind == MN_GRAD)
            mn_rows<real_t, MN_GRAD><<<rows_grid, 256, 0, st>>>(Zp, Rp, p.splits, zstride, (int) B, (int) K, wb, ldw, vb, Y, ldy, labels, sw,
                                                               need_out ? DT : nullptr, p.bpad, loss_dev ? terms : nullptr);
        else
            mn_rows<real_t, MN_HVP><<<rows_grid, 256, 0, st>>>(Zp, Rp, p.splits, zstride, (int) B, (int) K, wb, ldw, vb, Y, ldy, labels, sw,
                                                              DT, p.bpad, nullptr);
    }
    if (loss_dev && kind == MN_GRAD) {
        mn_wnorm<real_t><<<(unsigned) K, 256, 0, st>>>(w, ldw, (int) d, terms + B);
        mn_loss_finish<<<1, 256, 0, st>>>(terms, (int) B, terms + B, (int) K, (double) alpha, loss_dev);
        launched += 2;
    }
    if (need_out) {
        dim3 tb(32, 8), tg((unsigned) ((d + 31) / 32), (unsigned) ((B + 31) / 32));
        mn_transpose<real_t><<<tg, tb, 0, st>>>(X, ldx, XT, p.bpad, (int) B, (int) d);
        ++launched;
        if (int r = mn_check("multinomial rows", launched)) return r;
        // GEMM 2: G = DT XT'
        // (+ alpha * W or alpha * V fused into the tensor-core epilogue: saves a read-modify-write pass over the gradient)
        bool fused = false;
        if (int r = launch_gemm(DT, p.bpad, XT, p.bpad, out, ldw, 0, (int) K, (int) d, (int) B, 1, st,
                                kind == MN_HVP ? v : w, ldw, alpha, &fused, scatter)) return r;
        if (scatter && scatter->world > 1 && !fused) return -5;
        if (fused) {
            if (!fit_intercept) return 0;
            mn_intercept<real_t><<<(unsigned) ((K + 7) / 8), 256, 0, st>>>(out, ldw, (int) d, DT, p.bpad, (int) B, (int) K,
                                                                           scatter ? *scatter : MnScatter());
            return mn_check("multinomial intercept", 1);
        }
        mn_finish<real_t><<<(unsigned) K, 256, 0, st>>>(out, ldw, kind == MN_HVP ? v : w, ldw, (int) d, fit_intercept, alpha, DT, p.bpad, (int) B);
        return mn_check("multinomial finish", 1);
    }
    return mn_check("multinomial rows", launched);
}

}  // namespace

void stochqn_b200_internal_mn_small_plan(long long B, long long d, long long K, int grid, StochqnMnSmallPlan* out)
{
    out->ok = 0; out->off_zp = 0; out->off_dg = 0; out->smem = 0;
    if (B <= 0 || d <= 0 || K <= 0 || grid < 1 || grid > MS_MAX_GRID) return;
    const char* e = getenv("STOCHQN_B200_MN_SMALL");
    if (e && atoi(e) == 0) return;
    const MnPlan p = mn_plan(B, d, K);
    if (!p.small_ok || B > MS_MAX_B || K > MS_MAX_K || B * K > MS_MAX_BK || (d + grid - 1) / grid > MS_CW) return;
    out->smem = mn_small_smem(B, d, K, grid);
    if (out->smem > 200 * 1024) return;
    out->off_zp = p.off_small;
    out->off_dg = p.off_dt;
    out->ok = 1;
}

extern "C" {

int stochqn_b200_gemm_tn(const real_t* A, long long lda, const real_t* B, long long ldb, real_t* C, long long ldc,
                         int M, int N, int K, void* stream)
{
    if (!A || !B || !C || M <= 0 || N <= 0 || K <= 0) return -1;
    return launch_gemm(A, lda, B, ldb, C, ldc, 0, M, N, K, 1, (cudaStream_t) stream);
}

int stochqn_b200_debug_mn_trace(unsigned long long* dev_buf)
{
    g_mn_trace = dev_buf;
    return 0;
}

size_t stochqn_b200_multinomial_work_size(long long nrows, long long nfeat, long long nclasses)
{
    if (nrows <= 0 || nfeat <= 0 || nclasses <= 0) return 0;
    return mn_plan(nrows, nfeat, nclasses).total;
}

int stochqn_b200_multinomial_loss_grad(const real_t* X, long long ldx, const real_t* Y, long long ldy, const int* labels,
                                       const real_t* sw, long long nrows, long long nfeat, long long nclasses,
                                       int fit_intercept, const real_t* w, real_t alpha, real_t* grad, double* loss_dev,
                                       void* work, void* stream)
{
    if (!grad && !loss_dev) return -1;
    return mn_common(MN_GRAD, X, ldx, Y, ldy, labels, sw, nrows, nfeat, nclasses, fit_intercept, w, nullptr, alpha, grad, loss_dev,
                     work, (cudaStream_t) stream);
}

int stochqn_b200_multinomial_grad_reduce_scatter(void* comm, const real_t* X, long long ldx, const real_t* Y, long long ldy,
                                                 const int* labels, const real_t* sw, long long nrows, long long nfeat,
                                                 long long nclasses, int fit_intercept, const real_t* w, real_t alpha,
                                                 real_t* grad_block, long long block_count, void* work, void* stream)
{
#ifdef USE_FLOAT
    const long long ldw = nfeat + (fit_intercept ? 1 : 0);
    StochqnRsPlan plan;
    if (!grad_block || block_count < ldw) return -5;
    if (int r = stochqn_b200_internal_rs_begin(comm, block_count, &plan)) return r;       // -5: no peer-memory path
    if ((long long) plan.world * block_count < nclasses * ldw) return -1;
    MnScatter sc;
    sc.world = plan.world; sc.blk = block_count;
    for (int r = 0; r < plan.world; ++r) sc.dst[r] = plan.dst[r];
    cudaStream_t st = (cudaStream_t) stream;
    // `out` is only a flat index space here (leading dimension ldw); nothing is stored through it
    if (int r = mn_common(MN_GRAD, X, ldx, Y, ldy, labels, sw, nrows, nfeat, nclasses, fit_intercept, w, nullptr, alpha,
                          grad_block, nullptr, work, st, &sc)) return r;
    if (int r = stochqn_b200_internal_barrier(comm, st)) return r;          // every sender's tiles have landed
    const bool v4 = (block_count % 4 == 0) && ((((uintptr_t) grad_block) | ((uintptr_t) plan.local)) & 15u) == 0;
    if (v4) mn_rs_reduce<real_t, 4><<<1184, 256, 0, st>>>((const real_t*) plan.local, block_count, plan.world, grad_block);
    else    mn_rs_reduce<real_t, 1><<<1184, 256, 0, st>>>((const real_t*) plan.local, block_count, plan.world, grad_block);
    return mn_check("multinomial reduce-scatter", 1);
#else
    (void) comm; (void) X; (void) ldx; (void) Y; (void) ldy; (void) labels; (void) sw; (void) nrows; (void) nfeat; (void) nclasses;
    (void) fit_intercept; (void) w; (void) alpha; (void) grad_block; (void) block_count; (void) work; (void) stream;
    return -5;          // the fused path lives in the tensor-core (float) build
#endif
}

int stochqn_b200_multinomial_hess_vec(const real_t* X, long long ldx, const real_t* Y, long long ldy, const int* labels,
                                      const real_t* sw, long long nrows, long long nfeat, long long nclasses,
                                      int fit_intercept, const real_t* w, const real_t* v, real_t alpha, real_t* hess_vec,
                                      void* work, void* stream)
{
    if (!v || !hess_vec) return -1;
    return mn_common(MN_HVP, X, ldx, Y, ldy, labels, sw, nrows, nfeat, nclasses, fit_intercept, w, v, alpha, hess_vec, nullptr,
                     work, (cudaStream_t) stream);
}

}  // extern "C"
