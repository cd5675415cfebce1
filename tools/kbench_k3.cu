// kbench_k3.cu - developer micro-benchmark: variants of the K3 combine + update pass (fp64, m = 10, oLBFGS mode),
// timed alone with CUDA events.  Not part of the product; used to pick the work split and the row stride.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/kbench_k3 tools/kbench_k3.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <math.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int M = 10;

__device__ __forceinline__ double2 ld_nc(const double* p, bool noalloc)
{
    double2 r;
    if (noalloc) asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    else r = __ldg(reinterpret_cast<const double2*>(p));
    return r;
}
__device__ __forceinline__ double2 ld_rw(const double* p, bool noalloc)
{
    double2 r;
    if (noalloc) asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p) : "memory");
    else r = *reinterpret_cast<const double2*>(p);
    return r;
}
__device__ __forceinline__ void st_cs(double* p, double2 v, bool streaming)
{
    if (streaming) asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" :: "l"(p), "d"(v.x), "d"(v.y) : "memory");
    else *reinterpret_cast<double2*>(p) = v;
}

// ---------------------------------------------------------------------------------------------------------------
// A: every thread owns U chunks and streams ALL 2M rows for them (no shared memory, no barrier)
// ---------------------------------------------------------------------------------------------------------------
template <int THREADS, int U, bool NOALLOC, bool STCS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k3_allrows(const double* g, double* gout, double* S, const double* __restrict__ Y, size_t ld, int new_slot, long long n,
           double* __restrict__ x, double step, const double* __restrict__ coef)
{
    double ca[M], cb[M];
    #pragma unroll
    for (int j = 0; j < M; ++j) { ca[j] = coef[j]; cb[j] = coef[M + j]; }
    const double gamma = coef[2 * M], nstep = -step;
    const long long nchunks = n / 2;
    const long long stride = (long long) gridDim.x * THREADS;
    for (long long c0 = (long long) blockIdx.x * THREADS + threadIdx.x; c0 < nchunks; c0 += stride * U) {
        double2 gv[U], xv[U], sv[U][M], yv[U][M];
        #pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long c = c0 + u * stride;
            if (c < nchunks) {
                const size_t off = (size_t) c * 2;
                gv[u] = ld_rw(g + off, false);
                xv[u] = ld_rw(x + off, false);
                #pragma unroll
                for (int j = 0; j < M; ++j) { sv[u][j] = ld_rw(S + (size_t) j * ld + off, NOALLOC); yv[u][j] = ld_nc(Y + (size_t) j * ld + off, NOALLOC); }
            }
        }
        #pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long c = c0 + u * stride;
            if (c < nchunks) {
                const size_t off = (size_t) c * 2;
                double2 d;
                d.x = gamma * gv[u].x; d.y = gamma * gv[u].y;
                #pragma unroll
                for (int j = 0; j < M; ++j) {
                    d.x = fma(ca[j], sv[u][j].x, d.x); d.y = fma(ca[j], sv[u][j].y, d.y);
                    d.x = fma(cb[j], yv[u][j].x, d.x); d.y = fma(cb[j], yv[u][j].y, d.y);
                }
                double2 xn, sn;
                xn.x = fma(nstep, d.x, xv[u].x); xn.y = fma(nstep, d.y, xv[u].y);
                sn.x = nstep * d.x; sn.y = nstep * d.y;
                st_cs(x + off, xn, STCS);
                st_cs(S + (size_t) new_slot * ld + off, sn, STCS);
                st_cs(gout + off, sn, STCS);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// B: GROUPS row-groups x LANES chunk-lanes with a shared-memory exchange (the shipped structure), U chunks per barrier
// ---------------------------------------------------------------------------------------------------------------
template <int GROUPS, int LANES, int RPG, int U, bool NOALLOC, int MINB>
__global__ void __launch_bounds__(GROUPS * LANES, MINB)
k3_groups(const double* g, double* gout, double* S, const double* __restrict__ Y, size_t ld, int new_slot, long long n,
          double* __restrict__ x, double step, const double* __restrict__ coef)
{
    const int group = threadIdx.x / LANES, lane = threadIdx.x % LANES;
    const double* rows[RPG];
    double cf[RPG];
    #pragma unroll
    for (int r = 0; r < RPG; ++r) {
        int v = group * RPG + r;
        const bool live = v < 2 * M;
        if (!live) v = 2 * M - 1;
        rows[r] = v < M ? S + (size_t) v * ld : Y + (size_t) (v - M) * ld;
        cf[r] = live ? coef[v] : 0.0;
    }
    const double gamma = group == 0 ? coef[2 * M] : 0.0, nstep = -step;
    __shared__ __align__(16) double xchg[2][U][GROUPS][LANES * 2];
    int buf = 0;
    const long long nchunks = n / 2;
    const long long stride = (long long) gridDim.x * LANES;
    for (long long base = (long long) blockIdx.x * LANES; base < nchunks; base += stride * U) {
        double2 rv[U][RPG], gv[U], xv[U];
        #pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long c = base + u * stride + lane;
            if (c < nchunks) {
                const size_t off = (size_t) c * 2;
                #pragma unroll
                for (int r = 0; r < RPG; ++r) rv[u][r] = ld_rw(rows[r] + off, NOALLOC);
                if (group == 0) { gv[u] = ld_rw(g + off, false); xv[u] = ld_rw(x + off, false); }
            }
        }
        #pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long c = base + u * stride + lane;
            double2 part = make_double2(0.0, 0.0);
            if (c < nchunks) {
                if (group == 0) { part.x = gamma * gv[u].x; part.y = gamma * gv[u].y; }
                #pragma unroll
                for (int r = 0; r < RPG; ++r) { part.x = fma(cf[r], rv[u][r].x, part.x); part.y = fma(cf[r], rv[u][r].y, part.y); }
            }
            *reinterpret_cast<double2*>(&xchg[buf][u][group][lane * 2]) = part;
        }
        __syncthreads();
        #pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long c = base + u * stride + lane;
            if (c < nchunks && group < 3) {
                const size_t off = (size_t) c * 2;
                double2 d = *reinterpret_cast<double2*>(&xchg[buf][u][0][lane * 2]);
                #pragma unroll
                for (int q = 1; q < GROUPS; ++q) {
                    double2 t = *reinterpret_cast<double2*>(&xchg[buf][u][q][lane * 2]);
                    d.x += t.x; d.y += t.y;
                }
                if (group == 0) {
                    double2 xn;
                    xn.x = fma(nstep, d.x, xv[u].x); xn.y = fma(nstep, d.y, xv[u].y);
                    *reinterpret_cast<double2*>(x + off) = xn;
                } else {
                    double2 sn;
                    sn.x = nstep * d.x; sn.y = nstep * d.y;
                    if (group == 1) *reinterpret_cast<double2*>(S + (size_t) new_slot * ld + off) = sn;
                    if (group == GROUPS - 1 || group == 2) *reinterpret_cast<double2*>(gout + off) = sn;
                }
            }
        }
        buf ^= 1;
    }
}

__global__ void fill_kernel(double* p, long long n, unsigned seed, double scale)
{
    const long long stride = (long long) gridDim.x * blockDim.x;
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        unsigned h = (unsigned) (i * 2654435761u) ^ seed;
        h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
        p[i] = scale * ((double) (h & 0xffff) / 65536.0 - 0.5);
    }
}
__global__ void copy_kernel(const double2* __restrict__ a, double2* __restrict__ b, long long n2)
{
    const long long stride = (long long) gridDim.x * blockDim.x;
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) b[i] = a[i];
}

struct Bufs { double *g, *gout, *S, *Y, *x, *coef; size_t ld; long long n; };

template <typename F>
float time_it(F launch, int reps = 10)
{
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; ++i) launch();
    CK(cudaEventRecord(b));
    CK(cudaDeviceSynchronize());
    CK(cudaGetLastError());
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    return ms / reps;
}

void report(const char* name, float ms, const Bufs& B, int grid, int occ)
{
    const double bytes = 24.0 * B.n * 8;       // algorithmic: 22 reads + x + s_new (the grad write-back is extra, not counted)
    printf("%-52s occ %d grid %5d  %8.3f ms  %8.1f GB/s (algorithmic 24 vec; 25 moved)\n", name, occ, grid, ms, bytes / ms / 1e6);
    fflush(stdout);
}

template <int THREADS, int U, bool NOALLOC, bool STCS, int MINB>
void run_all(const char* name, const Bufs& B, int sms)
{
    auto kern = k3_allrows<THREADS, U, NOALLOC, STCS, MINB>;
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, 0));
    const int grid = sms * occ;
    float ms = time_it([&] { kern<<<grid, THREADS>>>(B.g, B.gout, B.S, B.Y, B.ld, 3, B.n, B.x, 1e-4, B.coef); });
    report(name, ms, B, grid, occ);
}

template <int GROUPS, int LANES, int RPG, int U, bool NOALLOC, int MINB>
void run_groups(const char* name, const Bufs& B, int sms)
{
    auto kern = k3_groups<GROUPS, LANES, RPG, U, NOALLOC, MINB>;
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, GROUPS * LANES, 0));
    const int grid = sms * occ;
    float ms = time_it([&] { kern<<<grid, GROUPS * LANES>>>(B.g, B.gout, B.S, B.Y, B.ld, 3, B.n, B.x, 1e-4, B.coef); });
    report(name, ms, B, grid, occ);
}

int main(int argc, char** argv)
{
    long long n = argc > 1 ? atoll(argv[1]) : (1ll << 26);
    long long pad = argc > 2 ? atoll(argv[2]) : 0;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    Bufs B;
    B.n = n; B.ld = (size_t) (n + pad);
    CK(cudaMalloc(&B.g, n * 8)); CK(cudaMalloc(&B.gout, n * 8)); CK(cudaMalloc(&B.x, n * 8));
    CK(cudaMalloc(&B.S, (size_t) M * B.ld * 8)); CK(cudaMalloc(&B.Y, (size_t) M * B.ld * 8));
    CK(cudaMalloc(&B.coef, 64 * 8));
    fill_kernel<<<sms * 8, 256>>>(B.g, n, 1u, 1.0);
    fill_kernel<<<sms * 8, 256>>>(B.x, n, 5u, 1.0);
    fill_kernel<<<sms * 8, 256>>>(B.S, (long long) M * B.ld, 2u, 1.0);
    fill_kernel<<<sms * 8, 256>>>(B.Y, (long long) M * B.ld, 3u, 1.0);
    fill_kernel<<<1, 64>>>(B.coef, 64, 7u, 1e-3);
    CK(cudaDeviceSynchronize());
    printf("device %s, %d SMs, n = %lld (vec %.0f MiB), row stride n + %lld elements\n", prop.name, sms, n, n * 8 / 1048576.0, pad);
    {
        float ms = time_it([&] { copy_kernel<<<sms * 8, 256>>>((const double2*) B.S, (double2*) B.Y, (long long) 4 * n / 2); });
        printf("%-52s                   %8.3f ms  %8.1f GB/s (read+write)\n", "copy 4 vec -> 4 vec (plain LDG/STG)", ms, 8.0 * n * 8 / ms / 1e6);
    }
    run_groups<4, 64, 5, 1, false, 1>("groups g4 l64 rpg5 u1 (shipped)", B, sms);
    run_groups<4, 64, 5, 1, true, 3>("groups g4 l64 rpg5 u1 noalloc minb3", B, sms);
    run_groups<4, 64, 5, 2, true, 2>("groups g4 l64 rpg5 u2 noalloc minb2", B, sms);
    run_groups<4, 64, 5, 2, false, 2>("groups g4 l64 rpg5 u2 minb2", B, sms);
    run_groups<4, 32, 5, 2, true, 4>("groups g4 l32 rpg5 u2 noalloc minb4 (128 thr)", B, sms);
    run_groups<2, 64, 10, 1, true, 4>("groups g2 l64 rpg10 u1 noalloc minb4 (128 thr)", B, sms);
    run_groups<2, 64, 10, 2, true, 2>("groups g2 l64 rpg10 u2 noalloc minb2 (128 thr)", B, sms);
    run_groups<2, 128, 10, 1, true, 2>("groups g2 l128 rpg10 u1 noalloc minb2", B, sms);
    run_groups<4, 64, 5, 2, false, 3>("groups g4 l64 rpg5 u2 minb3", B, sms);
    run_groups<2, 64, 10, 1, false, 4>("groups g2 l64 rpg10 u1 minb4 (128 thr)", B, sms);
    run_groups<2, 128, 10, 1, false, 2>("groups g2 l128 rpg10 u1 minb2", B, sms);
    run_groups<2, 32, 10, 1, true, 8>("groups g2 l32 rpg10 u1 noalloc minb8 (64 thr)", B, sms);
    if (argc > 3) return 0;
    run_all<128, 1, true, false, 2>("allrows 128thr u1 noalloc minb2", B, sms);
    run_all<128, 1, true, false, 3>("allrows 128thr u1 noalloc minb3", B, sms);
    run_all<128, 1, true, false, 4>("allrows 128thr u1 noalloc minb4", B, sms);
    run_all<128, 1, false, false, 4>("allrows 128thr u1 minb4", B, sms);
    run_all<128, 1, true, true, 4>("allrows 128thr u1 noalloc st.cs minb4", B, sms);
    run_all<128, 1, true, false, 6>("allrows 128thr u1 noalloc minb6", B, sms);
    run_all<256, 1, true, false, 1>("allrows 256thr u1 noalloc minb1", B, sms);
    run_all<256, 1, true, false, 2>("allrows 256thr u1 noalloc minb2", B, sms);
    run_all<256, 1, true, true, 2>("allrows 256thr u1 noalloc st.cs minb2", B, sms);
    run_all<256, 1, true, false, 3>("allrows 256thr u1 noalloc minb3", B, sms);
    run_all<128, 2, true, false, 2>("allrows 128thr u2 noalloc minb2", B, sms);
    run_all<64, 2, true, false, 4>("allrows 64thr u2 noalloc minb4", B, sms);
    run_all<64, 1, true, false, 8>("allrows 64thr u1 noalloc minb8", B, sms);
    run_all<512, 1, true, false, 1>("allrows 512thr u1 noalloc minb1", B, sms);
    return 0;
}
