// callbacks.cu - bundled device callbacks that serve the optimizers' requests without leaving
// the GPU: the chained Rosenbrock function of the reference's example programs
// (example/c_rosen.c:13-41) and the binary-logistic closed forms of the reference's R model
// (R/logistic.R:1-37).  C ABI declared in include/stochqn_b200.h.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include <atomic>

#include "stochqn.h"
#include "stochqn_b200.h"
#include "vecio.cuh"

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

// One 16-byte chunk per thread and iteration: the chunk is loaded as a vector, the two neighbours as
// scalars (they are the edge elements of the adjacent threads' chunks: L1 hits), the result stored as a vector.
template <int VEC>
__global__ void __launch_bounds__(kT)
rosen_grad_kernel(const real_t* __restrict__ x, real_t* __restrict__ g, long long n, long long offset,
                  long long n_global, const real_t* __restrict__ halo)
{
    const long long nchunks = n / VEC;
    const long long stride = (long long) gridDim.x * blockDim.x;
    for (long long c = (long long) blockIdx.x * blockDim.x + threadIdx.x; c < nchunks; c += stride) {
        const long long i0 = c * VEC;
        sqn::Pack<real_t, VEC> xv = sqn::ld_stream<real_t, VEC>(x + i0), gv;
        double v[VEC + 2];
        v[0] = (i0 > 0) ? (double) __ldg(x + i0 - 1) : (offset > 0 ? (double) halo[0] : 0.0);
        #pragma unroll
        for (int e = 0; e < VEC; ++e) v[e + 1] = (double) xv.get(e);
        v[VEC + 1] = (i0 + VEC < n) ? (double) __ldg(x + i0 + VEC) : (offset + n < n_global ? (double) halo[1] : 0.0);
        #pragma unroll
        for (int e = 0; e < VEC; ++e) {
            const long long gi = i0 + e + offset;
            gv.set(e, (real_t) rosen_g(v[e], v[e + 1], v[e + 2], gi > 0, gi < n_global - 1));
        }
        sqn::st_vec<real_t, VEC>(g + i0, gv);
    }
    if (VEC > 1 && blockIdx.x == 0) {                  // scalar tail
        const long long i = nchunks * VEC + threadIdx.x;
        if (i < n) {
            const long long gi = i + offset;
            const double xm = (i > 0) ? (double) x[i - 1] : (offset > 0 ? (double) halo[0] : 0.0);
            const double xp = (i < n - 1) ? (double) x[i + 1] : (offset + n < n_global ? (double) halo[1] : 0.0);
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

// ---- binary logistic regression (first version: two sweeps of the batch) ---------------------------
// pass A: one warp per row: z = x_row'w (and t = x_row'v), r_row written to scratch
//         grad: r = (sigmoid(z) - y) * sw          hess_vec: r = p(1-p) * sw * t        loss: per-row loss * sw
// pass B: one thread per column: out[col] = sum_rows r_row X[row][col] / sum(sw) + 2*lambda*u[col]
enum { LG_GRAD = 0, LG_HVP = 1, LG_LOSS = 2 };

template <int KIND>
__global__ void __launch_bounds__(kT)
logistic_rows_kernel(const real_t* __restrict__ X, long long ldx, const real_t* __restrict__ y,
                     const real_t* __restrict__ sw, long long nrows, long long ncols,
                     const real_t* __restrict__ w, const real_t* __restrict__ v, double* __restrict__ r)
{
    const int lane = threadIdx.x & 31;
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
            const double p = 1.0 / (1.0 + exp(-z));
            const double wt = sw ? (double) sw[row] : 1.0;
            if (KIND == LG_GRAD) r[row] = (p - (double) y[row]) * wt;
            else if (KIND == LG_HVP) r[row] = p * (1.0 - p) * wt * t;
            else { const double yy = (double) y[row]; r[row] = -(yy * log(p) + (1.0 - yy) * log(1.0 - p)) * wt; }
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
                       real_t* __restrict__ out)
{
    const long long c = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncols) return;
    double acc = 0;
    for (int s = 0; s < slices; ++s) acc += colpart[(long long) s * ncols + c];
    out[c] = (real_t) (acc / r[nrows] + 2.0 * (double) lambda * (double) u[c]);
}

__global__ void logistic_loss_finish_kernel(const double* __restrict__ r, long long nrows, const real_t* __restrict__ w,
                                            long long ncols, real_t lambda, double* __restrict__ loss)
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
        *loss = r[nrows + 1] / r[nrows] + (double) lambda * s;
    }
}

int row_slices(long long nrows)
{
    long long s = nrows / 64;
    if (s < 1) s = 1;
    if (s > 64) s = 64;
    return (int) s;
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
    constexpr int V = 16 / sizeof(real_t);
    if (((((uintptr_t) x) | ((uintptr_t) grad)) & 15u) == 0)
        rosen_grad_kernel<V><<<grid_1d(n_local / V), kT, 0, (cudaStream_t) stream>>>(x, grad, n_local, offset, n_global, halo);
    else
        rosen_grad_kernel<1><<<grid_1d(n_local), kT, 0, (cudaStream_t) stream>>>(x, grad, n_local, offset, n_global, halo);
    return check_launch("rosenbrock_grad");
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

size_t stochqn_b200_logistic_work_size(long long nrows, long long ncols)
{
    return sizeof(double) * (size_t) (nrows + 2 + (long long) row_slices(nrows) * ncols);
}

static int logistic_common(int kind, const real_t* X, long long ldx, const real_t* y, const real_t* sw,
                           long long nrows, long long ncols, const real_t* w, const real_t* v, real_t lambda,
                           real_t* out, double* loss, void* work, cudaStream_t st)
{
    double* r = (double*) work;
    double* colpart = r + nrows + 2;
    const int g_rows = grid_1d(nrows * 32);
    if (kind == LG_GRAD) logistic_rows_kernel<LG_GRAD><<<g_rows, kT, 0, st>>>(X, ldx, y, sw, nrows, ncols, w, v, r);
    else if (kind == LG_HVP) logistic_rows_kernel<LG_HVP><<<g_rows, kT, 0, st>>>(X, ldx, y, sw, nrows, ncols, w, v, r);
    else logistic_rows_kernel<LG_LOSS><<<g_rows, kT, 0, st>>>(X, ldx, y, sw, nrows, ncols, w, v, r);
    logistic_norm_kernel<<<1, kT, 0, st>>>(sw, nrows, r);
    if (kind == LG_LOSS) {
        logistic_loss_finish_kernel<<<1, kT, 0, st>>>(r, nrows, w, ncols, lambda, loss);
    } else {
        const int slices = row_slices(nrows);
        dim3 grid((unsigned) ((ncols + kT - 1) / kT), (unsigned) slices);
        logistic_cols_kernel<<<grid, kT, 0, st>>>(X, ldx, nrows, ncols, r, colpart);
        logistic_finish_kernel<<<(unsigned) ((ncols + kT - 1) / kT), kT, 0, st>>>(colpart, slices, nrows, ncols, r,
                                                                                  kind == LG_HVP ? v : w, lambda, out);
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

}  // extern "C"
