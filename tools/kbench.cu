// kbench.cu - developer micro-benchmark: variants of the K1 multi-dot pass (fp64, m = 10, pending pair),
// timed alone with CUDA events at n = 2^26.  Not part of the product; used to pick the work split.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o kbench tools/kbench.cu
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <math.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int M = 10;
constexpr int NSUM = 4 * M + 2;

__device__ __forceinline__ double warp_sum(double v)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double2 ldg_stream(const double* p, bool noalloc)
{
    double2 r;
    if (noalloc) asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    else r = __ldg(reinterpret_cast<const double2*>(p));
    return r;
}

// ---------------------------------------------------------------------------------------------------------------
// LDG variant: GROUPS row-groups x LANES chunk-lanes, RPG virtual rows per group, UNROLL chunks per thread in flight
// ---------------------------------------------------------------------------------------------------------------
template <int GROUPS, int LANES, int RPG, int UNROLL, bool NOALLOC, int MINB>
__global__ void __launch_bounds__(GROUPS * LANES, MINB)
k1x(const double* __restrict__ g, const double* __restrict__ S, const double* __restrict__ Y, size_t ld, int pend,
    long long n, double* __restrict__ grad_prev, double* __restrict__ partials)
{
    constexpr int T = GROUPS * LANES;
    const int group = threadIdx.x / LANES, lane = threadIdx.x % LANES;
    const bool lead = group == 0;
    if (group != GROUPS - 1) grad_prev = nullptr;
    const double* sc_row = S + (size_t) pend * ld;
    const double* yc_row = Y + (size_t) pend * ld;
    const double* rows[RPG];
    #pragma unroll
    for (int r = 0; r < RPG; ++r) {
        int v = group * RPG + r;
        if (v >= 2 * M) v = 2 * M - 1;
        rows[r] = v < M ? S + (size_t) v * ld : Y + (size_t) (v - M) * ld;
    }
    double a_g[RPG], a_c[RPG], a_gg = 0, a_ss = 0;
    #pragma unroll
    for (int r = 0; r < RPG; ++r) { a_g[r] = 0; a_c[r] = 0; }
    const long long nchunks = n / 2;
    const long long stride = (long long) gridDim.x * LANES;
    for (long long c0 = (long long) blockIdx.x * LANES + lane; c0 < nchunks; c0 += stride * UNROLL) {
        double2 gv[UNROLL], yc[UNROLL], sc[UNROLL], rv[UNROLL][RPG];
        #pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const long long c = c0 + u * stride;
            if (c < nchunks) {
                const size_t off = (size_t) c * 2;
                gv[u] = ldg_stream(g + off, false);
                yc[u] = ldg_stream(yc_row + off, false);
                if (lead) sc[u] = ldg_stream(sc_row + off, false);
                #pragma unroll
                for (int r = 0; r < RPG; ++r) rv[u][r] = ldg_stream(rows[r] + off, NOALLOC);
            }
        }
        #pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const long long c = c0 + u * stride;
            if (c < nchunks) {
                if (grad_prev) *reinterpret_cast<double2*>(grad_prev + (size_t) c * 2) = gv[u];
                if (lead) {
                    a_gg = fma(gv[u].x, gv[u].x, a_gg); a_gg = fma(gv[u].y, gv[u].y, a_gg);
                    a_ss = fma(sc[u].x, sc[u].x, a_ss); a_ss = fma(sc[u].y, sc[u].y, a_ss);
                }
                #pragma unroll
                for (int r = 0; r < RPG; ++r) {
                    a_g[r] = fma(rv[u][r].x, gv[u].x, a_g[r]); a_g[r] = fma(rv[u][r].y, gv[u].y, a_g[r]);
                    a_c[r] = fma(rv[u][r].x, yc[u].x, a_c[r]); a_c[r] = fma(rv[u][r].y, yc[u].y, a_c[r]);
                }
            }
        }
    }
    constexpr int NA = 2 * RPG + 2, W = T / 32, WPG = LANES / 32;
    __shared__ double red[W][NA];
    const int warp = threadIdx.x >> 5, wl = threadIdx.x & 31;
    #pragma unroll
    for (int p = 0; p < NA; ++p) {
        double v = p < RPG ? a_g[p < RPG ? p : 0] : p < 2 * RPG ? a_c[(p - RPG) < RPG ? (p - RPG) : 0] : p == 2 * RPG ? a_gg : a_ss;
        v = warp_sum(v);
        if (wl == 0) red[warp][p] = v;
    }
    __syncthreads();
    double* out = partials + (size_t) blockIdx.x * NSUM;
    for (int t = threadIdx.x; t < GROUPS * NA; t += T) {
        const int gi = t / NA, p = t % NA;
        double v = 0;
        for (int w = 0; w < WPG; ++w) v += red[gi * WPG + w][p];
        if (p >= 2 * RPG) { if (gi == 0) out[4 * M + (p - 2 * RPG)] = v; continue; }
        const int r = p % RPG, vrow = gi * RPG + r;
        if (vrow >= 2 * M) continue;
        const bool is_s = vrow < M;
        const int j = is_s ? vrow : vrow - M;
        const int k = (p < RPG) ? (is_s ? 0 : 1) : (is_s ? 2 : 3);
        out[k * M + j] = v;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// TMA variant: 1-D bulk copies (cp.async.bulk) of TE-element tiles of the 2M+1 rows into a STAGES-deep shared-memory
// ring, mbarrier full/empty handshake, one producer thread, 256 consumer threads (4 row-groups x 64 lanes).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok = 0;
    const uint32_t addr = smem_u32(bar);
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int TE, int STAGES>
__global__ void __launch_bounds__(288, 1)
k1_tma(const double* __restrict__ g, const double* __restrict__ S, const double* __restrict__ Y, size_t ld, int pend,
       long long n, double* __restrict__ grad_prev, double* __restrict__ partials)
{
    constexpr int NR = 2 * M + 1;                     // rows per stage: g, S[0..M), Y[0..M)
    constexpr int CONS = 256, GROUPS = 4, LANES = 64, RPG = 5;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* tiles = reinterpret_cast<double*>(smem_raw);                   // [STAGES][NR][TE]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t) STAGES * NR * TE * 8);
    uint64_t* empty = full + STAGES;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CONS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long ntiles = n / TE;                  // bench: n is a multiple of TE
    if (tid >= CONS) {
        // ---------------- producer warp: one elected lane streams the tiles ----------------
        if (tid == CONS) {
            int s = 0; uint32_t ph = 0;
            for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
                mbar_wait(&empty[s], ph ^ 1);
                mbar_expect_tx(&full[s], NR * TE * 8);
                double* dst = tiles + (size_t) s * NR * TE;
                const size_t off = (size_t) t * TE;
                bulk_g2s(dst, g + off, TE * 8, &full[s]);
                #pragma unroll 1
                for (int j = 0; j < M; ++j) {
                    bulk_g2s(dst + (size_t) (1 + j) * TE, S + (size_t) j * ld + off, TE * 8, &full[s]);
                    bulk_g2s(dst + (size_t) (1 + M + j) * TE, Y + (size_t) j * ld + off, TE * 8, &full[s]);
                }
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
        }
        return;
    }
    // ---------------- consumers ----------------
    const int group = tid / LANES, lane = tid % LANES;
    double a_g[RPG], a_c[RPG], a_gg = 0, a_ss = 0;
    #pragma unroll
    for (int r = 0; r < RPG; ++r) { a_g[r] = 0; a_c[r] = 0; }
    int s = 0; uint32_t ph = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        mbar_wait(&full[s], ph);
        const double* tile = tiles + (size_t) s * NR * TE;
        const double* gt = tile;
        const double* yct = tile + (size_t) (1 + M + pend) * TE;
        const double* sct = tile + (size_t) (1 + pend) * TE;
        #pragma unroll
        for (int e0 = 0; e0 < TE; e0 += 2 * LANES) {
            const int e = e0 + 2 * lane;
            const double2 gv = *reinterpret_cast<const double2*>(gt + e);
            const double2 yc = *reinterpret_cast<const double2*>(yct + e);
            if (group == 0) {
                const double2 sc = *reinterpret_cast<const double2*>(sct + e);
                a_gg = fma(gv.x, gv.x, a_gg); a_gg = fma(gv.y, gv.y, a_gg);
                a_ss = fma(sc.x, sc.x, a_ss); a_ss = fma(sc.y, sc.y, a_ss);
            }
            if (group == GROUPS - 1 && grad_prev) *reinterpret_cast<double2*>(grad_prev + (size_t) t * TE + e) = gv;
            #pragma unroll
            for (int r = 0; r < RPG; ++r) {
                const double2 rv = *reinterpret_cast<const double2*>(tile + (size_t) (1 + group * RPG + r) * TE + e);
                a_g[r] = fma(rv.x, gv.x, a_g[r]); a_g[r] = fma(rv.y, gv.y, a_g[r]);
                a_c[r] = fma(rv.x, yc.x, a_c[r]); a_c[r] = fma(rv.y, yc.y, a_c[r]);
            }
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&empty[s]);
        if (++s == STAGES) { s = 0; ph ^= 1; }
    }
    constexpr int NA = 2 * RPG + 2;
    __shared__ double red[8][NA];
    const int warp = tid >> 5, wl = tid & 31;
    #pragma unroll
    for (int p = 0; p < NA; ++p) {
        double v = p < RPG ? a_g[p < RPG ? p : 0] : p < 2 * RPG ? a_c[(p - RPG) < RPG ? (p - RPG) : 0] : p == 2 * RPG ? a_gg : a_ss;
        v = warp_sum(v);
        if (wl == 0) red[warp][p] = v;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");     // consumers only (the producer warp has left)
    double* out = partials + (size_t) blockIdx.x * NSUM;
    for (int t2 = tid; t2 < GROUPS * NA; t2 += CONS) {
        const int gi = t2 / NA, p = t2 % NA;
        double v = red[gi * 2][p] + red[gi * 2 + 1][p];
        if (p >= 2 * RPG) { if (gi == 0) out[4 * M + (p - 2 * RPG)] = v; continue; }
        const int r = p % RPG, vrow = gi * RPG + r;
        const bool is_s = vrow < M;
        const int j = is_s ? vrow : vrow - M;
        const int k = (p < RPG) ? (is_s ? 0 : 1) : (is_s ? 2 : 3);
        out[k * M + j] = v;
    }
}

// reference streaming copy for the denominator on this box
__global__ void copy_kernel(const double2* __restrict__ a, double2* __restrict__ b, long long n2)
{
    const long long stride = (long long) gridDim.x * blockDim.x;
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) b[i] = a[i];
}

__global__ void fill_kernel(double* p, long long n, unsigned seed)
{
    const long long stride = (long long) gridDim.x * blockDim.x;
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        unsigned h = (unsigned) (i * 2654435761u) ^ seed;
        h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
        p[i] = (double) (h & 0xffff) / 65536.0 - 0.5;
    }
}

struct Bufs { double *g, *S, *Y, *gp, *partials; size_t ld; long long n; };

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

std::vector<double> sums_of(const Bufs& B, int grid)
{
    std::vector<double> h((size_t) grid * NSUM), s(NSUM, 0.0);
    CK(cudaMemcpy(h.data(), B.partials, h.size() * 8, cudaMemcpyDeviceToHost));
    for (int b = 0; b < grid; ++b) for (int p = 0; p < NSUM; ++p) s[p] += h[(size_t) b * NSUM + p];
    return s;
}

std::vector<double> g_ref;
void report(const char* name, float ms, const Bufs& B, int grid)
{
    const double bytes = 22.0 * B.n * 8;
    auto s = sums_of(B, grid);
    double worst = 0;
    if (g_ref.empty()) g_ref = s;
    for (int p = 0; p < NSUM; ++p) worst = fmax(worst, fabs(s[p] - g_ref[p]) / fmax(fabs(g_ref[p]), 1e-300));
    printf("%-44s grid %5d  %8.3f ms  %8.1f GB/s   max rel diff vs first %.1e\n", name, grid, ms, bytes / ms / 1e6, worst);
    fflush(stdout);
}

template <int GROUPS, int LANES, int RPG, int UNROLL, bool NOALLOC, int MINB>
void run_ldg(const char* name, const Bufs& B, int sms)
{
    auto kern = k1x<GROUPS, LANES, RPG, UNROLL, NOALLOC, MINB>;
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, GROUPS * LANES, 0));
    const int grid = sms * occ;
    CK(cudaMemset(B.partials, 0, (size_t) 4096 * NSUM * 8));
    float ms = time_it([&] { kern<<<grid, GROUPS * LANES>>>(B.g, B.S, B.Y, B.ld, 3, B.n, B.gp, B.partials); });
    char buf[128];
    snprintf(buf, sizeof buf, "%s (occ %d)", name, occ);
    report(buf, ms, B, grid);
}

template <int TE, int STAGES>
void run_tma(const char* name, const Bufs& B, int sms)
{
    auto kern = k1_tma<TE, STAGES>;
    const size_t smem = (size_t) STAGES * (2 * M + 1) * TE * 8 + 2 * STAGES * 8;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 288, smem));
    const int grid = sms * occ;
    CK(cudaMemset(B.partials, 0, (size_t) 4096 * NSUM * 8));
    float ms = time_it([&] { kern<<<grid, 288, smem>>>(B.g, B.S, B.Y, B.ld, 3, B.n, B.gp, B.partials); });
    char buf[128];
    snprintf(buf, sizeof buf, "%s (occ %d, smem %zu KB)", name, occ, smem / 1024);
    report(buf, ms, B, grid);
}

int main(int argc, char** argv)
{
    long long n = argc > 1 ? atoll(argv[1]) : (1ll << 26);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    Bufs B;
    const long long pad = argc > 2 ? atoll(argv[2]) : 0;       // extra elements per row (row stride experiments)
    B.n = n; B.ld = (size_t) (n + pad);
    CK(cudaMalloc(&B.g, n * 8)); CK(cudaMalloc(&B.gp, n * 8));
    CK(cudaMalloc(&B.S, (size_t) M * B.ld * 8)); CK(cudaMalloc(&B.Y, (size_t) M * B.ld * 8));
    CK(cudaMalloc(&B.partials, (size_t) 4096 * NSUM * 8));
    fill_kernel<<<sms * 8, 256>>>(B.g, n, 1u);
    fill_kernel<<<sms * 8, 256>>>(B.S, (long long) M * B.ld, 2u);
    fill_kernel<<<sms * 8, 256>>>(B.Y, (long long) M * B.ld, 3u);
    CK(cudaDeviceSynchronize());
    printf("device %s, %d SMs, n = %lld (vec %.0f MiB), row stride n + %lld\n", prop.name, sms, n, n * 8 / 1048576.0, pad);
    {
        float ms = time_it([&] { copy_kernel<<<sms * 8, 256>>>((const double2*) B.S, (double2*) B.Y, (long long) 4 * n / 2); });
        printf("%-44s             %8.3f ms  %8.1f GB/s (read+write)\n", "copy 4 vec -> 4 vec (plain LDG/STG)", ms, 8.0 * n * 8 / ms / 1e6);
        fill_kernel<<<sms * 8, 256>>>(B.Y, (long long) M * B.ld, 3u);
        CK(cudaDeviceSynchronize());
    }
    run_ldg<4, 64, 5, 1, false, 3>("ldg g4 l64 rpg5 u1 minb3", B, sms);
    run_ldg<4, 64, 5, 2, true, 2>("ldg g4 l64 rpg5 u2 noalloc minb2", B, sms);
    run_ldg<4, 64, 5, 2, false, 2>("ldg g4 l64 rpg5 u2 minb2", B, sms);
    run_ldg<4, 64, 5, 2, true, 3>("ldg g4 l64 rpg5 u2 noalloc minb3", B, sms);
    run_ldg<4, 64, 5, 2, false, 3>("ldg g4 l64 rpg5 u2 minb3", B, sms);
    run_ldg<4, 64, 5, 3, true, 2>("ldg g4 l64 rpg5 u3 noalloc minb2", B, sms);
    run_ldg<4, 64, 5, 3, false, 2>("ldg g4 l64 rpg5 u3 minb2", B, sms);
    run_ldg<4, 64, 5, 4, true, 1>("ldg g4 l64 rpg5 u4 noalloc minb1", B, sms);
    run_ldg<4, 64, 5, 4, true, 2>("ldg g4 l64 rpg5 u4 noalloc minb2", B, sms);
    run_ldg<4, 32, 5, 2, true, 4>("ldg g4 l32 rpg5 u2 noalloc minb4 (128 thr)", B, sms);
    run_ldg<4, 32, 5, 2, true, 6>("ldg g4 l32 rpg5 u2 noalloc minb6 (128 thr)", B, sms);
    run_ldg<4, 32, 5, 3, true, 4>("ldg g4 l32 rpg5 u3 noalloc minb4 (128 thr)", B, sms);
    run_ldg<4, 32, 5, 4, true, 4>("ldg g4 l32 rpg5 u4 noalloc minb4 (128 thr)", B, sms);
    run_ldg<4, 128, 5, 2, true, 1>("ldg g4 l128 rpg5 u2 noalloc minb1 (512 thr)", B, sms);
    run_ldg<2, 64, 10, 2, true, 2>("ldg g2 l64 rpg10 u2 noalloc minb2 (128 thr)", B, sms);
    run_ldg<2, 128, 10, 2, true, 1>("ldg g2 l128 rpg10 u2 noalloc minb1", B, sms);
    run_ldg<1, 128, 20, 1, true, 2>("ldg g1 l128 rpg20 u1 noalloc minb2 (all rows/thread)", B, sms);
    run_tma<512, 2>("tma te512 st2", B, sms);
    run_tma<384, 3>("tma te384 st3", B, sms);
    run_tma<256, 4>("tma te256 st4", B, sms);
    run_tma<256, 3>("tma te256 st3", B, sms);
    return 0;
}
