// gemm_tf32_sm100.cuh - C[M x N] = A[M x K] * B[N x K]'  (fp32 storage, both operands K-contiguous, "TN") on the
// 5th-generation tensor cores of sm_100a: tcgen05.mma kind::tf32, operands staged in shared memory by TMA
// (cp.async.bulk.tensor, 128-byte swizzle), accumulators in tensor memory, epilogue through tcgen05.ld.
//
// This is the one compute-bound piece of the hot path: the two products of the multinomial-logistic gradient /
// Hessian-vector callbacks (multinomial.cu: Z = X W' and G = D' X, 2*B*d*K flop each; BASELINE config 5: 8192 features x
// 4096 classes, fp32).  Inputs are rounded to tf32 (10-bit mantissa) by the tensor core, products accumulate in fp32:
// relative error of a dot product ~ 2^-11 / sqrt(K) typical, <= 2^-10 worst case (stated tolerance; the fp64 build and
// STOCHQN_B200_NO_TENSOR_CORES=1 use the CUDA-core kernel).
//
// Structure (one persistent CTA per SM, 256 threads, warp-specialised; static round-robin over 128 x 128 output tiles):
//   warp 0, one lane   TMA producer: for every k-block (32 tf32 = 128 bytes per row) waits for a free stage, arms the
//                      stage's mbarrier with the byte count and issues two tensor copies (A: 128 rows, B: 128 rows)
//   warp 1, one lane   MMA issuer: waits for the stage, issues 4 x tcgen05.mma (M 128, N 128, K 8) into the
//                      accumulator stage in TMEM, tcgen05.commit -> frees the smem stage; after the last k-block
//                      tcgen05.commit -> tells the epilogue the accumulator is complete
//   warp 2             allocates / frees the 256 TMEM columns (2 accumulator stages x 128 fp32 columns)
//   warps 4-7          epilogue: tcgen05.ld (32 lanes x 32 columns per instruction) -> registers -> per-warp 32 x 32 transpose in
//                      shared memory -> global (row-major C, any ldc, bounds-masked, 128 contiguous bytes per store), then
//                      hand the accumulator stage back to the MMA warp
// Rows / columns / k beyond the matrices are zero-filled by TMA (no padding required of the caller); A, B need 16-byte
// aligned base pointers and leading dimensions that are multiples of 4 elements.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <stdint.h>
#include <stdlib.h>
#include <mutex>

namespace tf32gemm {

constexpr int BM = 128, BK = 32;                       // BK tf32 = 128 bytes = one swizzle atom
constexpr int UMMA_K = 8;                              // 32 bytes of K per tcgen05.mma for tf32
constexpr int ACC_STAGES = 2;
constexpr int THREADS = 256;
constexpr uint32_t A_BYTES = BM * BK * 4;
constexpr int EPI_LD = 33;                             // padded row of the epilogue staging tile (bank-conflict free)
constexpr size_t EPI_BYTES = (size_t) 4 * 32 * EPI_LD * 4;        // one 32 x 32 fp32 tile per epilogue warp
// Tile width BN = 128 (6 stages of 32 KB, 256 TMEM columns) or 256 (4 stages of 48 KB, all 512 TMEM columns): the wide
// tile moves 25 % fewer operand bytes from L2 per flop - the 128-wide kernel is L2-bound (ncu: 11 TB/s into the SMs).
template <int BN> struct Cfg {
    static constexpr int STAGES = BN == 128 ? 6 : 4;
    static constexpr uint32_t B_BYTES = BN * BK * 4, STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr size_t SMEM_BYTES = (size_t) STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers */ + EPI_BYTES;
    static constexpr uint32_t TMEM_COLS = ACC_STAGES * BN;        // a power of two >= 32
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
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
    const uint32_t addr = smem_u32(bar);
    uint32_t ok = 0;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// shared-memory matrix descriptor, K-major operand, 128-byte swizzle: rows 128 bytes apart inside an 8-row group,
// groups 1024 bytes apart (SBO), version 1 (sm_100), layout type 2 = SWIZZLE_128B  (cute/arch/mma_sm100_desc.hpp)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t) ((smem_addr >> 4) & 0x3fffu);                 // start address, bits [0,14)
    d |= (uint64_t) 1 << 16;                                      // leading byte offset (unused for swizzled K-major), bits [16,30)
    d |= (uint64_t) ((1024u >> 4) & 0x3fffu) << 32;               // stride byte offset, bits [32,46)
    d |= (uint64_t) 1 << 46;                                      // version, bits [46,48)
    d |= (uint64_t) 2 << 61;                                      // layout type, bits [61,64)
    return d;
}
// instruction descriptor: D fp32, A / B tf32, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int bn)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t) (bn >> 3) << 17) | ((uint32_t) (BM >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                 :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Fused reduce-scatter: when world > 1 the epilogue does not store into C; the result is viewed as the flat vector
// e = row*ldc + col, cut into `world` blocks of `blk` elements, and every element is written straight into the receive
// slot of the rank that owns its block (dst[o]: rank o's slot for THIS sender, a cudaIpc-mapped peer pointer - the
// stores travel over NVLink while the tensor pipe works on the next tile).  blk >= ldc, so a row crosses at most one
// block boundary.
struct ScatterArgs {
    int world = 0;
    float* dst[16] = {};
    long long blk = 0;
};

template <int BN>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 float* __restrict__ C, long long ldc, int M, int N, int K,
                 const float* __restrict__ addend, long long ld_add, float alpha, const __grid_constant__ ScatterArgs sc)
{
    constexpr int STAGES = Cfg<BN>::STAGES;
    constexpr uint32_t STAGE_BYTES = Cfg<BN>::STAGE_BYTES, TMEM_COLS = Cfg<BN>::TMEM_COLS;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t) 1023);
    unsigned char* tiles = smem;                                              // [STAGES][A 16 KB | B 16 KB], 1024-aligned
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t) STAGES * STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* acc_full = empty + STAGES;
    uint64_t* acc_empty = acc_full + ACC_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + ACC_STAGES);
    float* epi = reinterpret_cast<float*>(smem + (size_t) STAGES * STAGE_BYTES + 256);      // [4 warps][32][EPI_LD]

    __shared__ float* sc_ptr[4][32];                      // scatter-mode row tables of the four epilogue warps
    __shared__ float* sc_ptr2[4][32];
    __shared__ int sc_lim[4][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN;
    const int ntiles = tiles_m * tiles_n;
    const int kblocks = (K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < ACC_STAGES; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ================= TMA producer =================
            int s = 0; uint32_t ph = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                const int m0 = (t / tiles_n) * BM, n0 = (t % tiles_n) * BN;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&empty[s], ph ^ 1u);
                    mbar_expect_tx(&full[s], STAGE_BYTES);
                    unsigned char* a_dst = tiles + (size_t) s * STAGE_BYTES;
                    tma_load_2d(a_dst, &map_a, &full[s], kb * BK, m0);
                    tma_load_2d(a_dst + A_BYTES, &map_b, &full[s], kb * BK, n0);
                    if (++s == STAGES) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ================= MMA issuer =================
            constexpr uint32_t idesc = make_idesc(BN);
            int s = 0; uint32_t ph = 0;
            int as = 0; uint32_t aph = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                mbar_wait(&acc_empty[as], aph ^ 1u);                 // the epilogue has drained this accumulator stage
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t) (as * BN);
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(tiles + (size_t) s * STAGE_BYTES);
                    const uint32_t b_addr = a_addr + A_BYTES;
                    #pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t da = make_desc(a_addr + (uint32_t) (k * UMMA_K * 4));
                        const uint64_t db = make_desc(b_addr + (uint32_t) (k * UMMA_K * 4));
                        mma_tf32(tmem_d, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty[s]);                          // smem stage is free once these MMAs have read it
                    if (kb == kblocks - 1) umma_commit(&acc_full[as]);
                    if (++s == STAGES) { s = 0; ph ^= 1u; }
                }
                if (++as == ACC_STAGES) { as = 0; aph ^= 1u; }
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue: TMEM -> registers -> global =================
        const int q = warp & 3;                                      // TMEM lane quadrant this warp may read
        int as = 0; uint32_t aph = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
            const int m0 = (t / tiles_n) * BM, n0 = (t % tiles_n) * BN;
            mbar_wait(&acc_full[as], aph);
            tc_fence_after();
            // A lane owns one accumulator row in TMEM; storing from there would scatter every store over 32 rows.
            // Each warp transposes its 32 x 32 block through shared memory so that a store instruction writes 32
            // consecutive floats of ONE row of C (128 contiguous bytes, whatever ldc is).
            float* stage = epi + (size_t) q * 32 * EPI_LD;
            // scatter mode: lane r works out, once per tile, where row (m0 + 32q + r) of this tile lands: a pointer
            // into its owner's receive slot for column n0, how many columns that owner still takes, and the start of the
            // next owner's slot (a row crosses at most one block boundary).  The table is read back as broadcasts.
            if (sc.world > 1) {
                const long long e = (long long) (m0 + q * 32 + lane) * ldc + n0;
                const int o = (int) (e / sc.blk);
                const long long rem = e - (long long) o * sc.blk;
                const long long lim = sc.blk - rem;
                sc_ptr[q][lane] = o < sc.world ? sc.dst[o] + rem : nullptr;
                sc_ptr2[q][lane] = o + 1 < sc.world ? sc.dst[o + 1] : nullptr;
                sc_lim[q][lane] = lim > (1ll << 30) ? (1 << 30) : (int) lim;
                __syncwarp();
            }
            #pragma unroll 1
            for (int c = 0; c < BN; c += 32) {
                uint32_t v[32];
                // fused epilogue C = A B' + alpha * addend: the 32 addend values of this lane's column are requested
                // up front (32 independent loads in flight, overlapping the TMEM read) - loading them one by one next
                // to the stores made the epilogue latency-bound and doubled the kernel time
                float ad[32];
                const int col = n0 + c + lane;
                if (addend) {
                    #pragma unroll
                    for (int r = 0; r < 32; ++r) {
                        const int row = m0 + q * 32 + r;
                        ad[r] = (row < M && col < N) ? __ldg(addend + (long long) row * ld_add + col) : 0.f;
                    }
                }
                const uint32_t taddr = tmem_base + ((uint32_t) (q * 32) << 16) + (uint32_t) (as * BN + c);
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                             "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                             "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                               "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                               "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                               "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                             : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                #pragma unroll
                for (int j = 0; j < 32; ++j) stage[lane * EPI_LD + j] = __uint_as_float(v[j]);
                __syncwarp();
                if (sc.world > 1) {
                    const int cc = c + lane;                        // column relative to n0
                    #pragma unroll 8
                    for (int r = 0; r < 32; ++r) {
                        const int row = m0 + q * 32 + r;
                        const int lim = sc_lim[q][r];
                        float* p = cc < lim ? sc_ptr[q][r] + cc : sc_ptr2[q][r] + (cc - lim);
                        float* base = cc < lim ? sc_ptr[q][r] : sc_ptr2[q][r];
                        if (row < M && col < N && base) {
                            float val = stage[r * EPI_LD + lane];
                            if (addend) val = fmaf(alpha, ad[r], val);
                            *p = val;
                        }
                    }
                } else if (addend) {
                    #pragma unroll
                    for (int r = 0; r < 32; ++r) {
                        const int row = m0 + q * 32 + r;
                        if (row < M && col < N) C[(long long) row * ldc + col] = fmaf(alpha, ad[r], stage[r * EPI_LD + lane]);
                    }
                } else {
                    #pragma unroll 4
                    for (int r = 0; r < 32; ++r) {
                        const int row = m0 + q * 32 + r;
                        if (row < M && col < N) C[(long long) row * ldc + col] = stage[r * EPI_LD + lane];
                    }
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[as]);
            if (++as == ACC_STAGES) { as = 0; aph ^= 1u; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// =====================================================================================================================
// The same product on CTA PAIRS (cta_group::2): two CTAs of a cluster on one TPC work on one 256 x 256 output tile.  Each
// CTA stages ITS 128 rows of A and ITS 128 rows of B (one half of the tile's N) per k-block - 32 KB instead of 48 KB for the
// same 128 x 256 share of the output, so a third less operand traffic from L2 (what limits the one-CTA kernel at 76 %
// tensor-pipe activity) and six pipeline stages instead of four.  The leader CTA (cluster rank 0) issues
// tcgen05.mma.cta_group::2 (M 256, N 256, K 8): every SM's tensor core multiplies its own A rows by both halves of B, read from
// both CTAs' shared memory, into its own 128 lanes x 256 columns of tensor memory.
//   full[s]       lives in the LEADER: both CTAs' TMA copies complete_tx on it (the peer addresses it with mapa)
//   empty[s]      one per CTA, released by tcgen05.commit.multicast (mask 0b11) once the MMAs have read the stage
//   acc_full[a]   one per CTA, same multicast commit after the last k-block
//   acc_empty[a]  lives in the leader, 8 arrivals: the four epilogue warps of each CTA
// =====================================================================================================================
constexpr int BN2 = 256;
constexpr int STAGES2 = 6;
constexpr uint32_t STAGE2_BYTES = 2 * A_BYTES;                         // A half + B half per CTA and stage
constexpr int THREADS2 = 384;                                            // warps 0-3: TMA, MMA, TMEM allocation, idle; warps 4-11: epilogue
constexpr int EPI_WARPS2 = 8;                                            // two warps per TMEM lane quadrant, alternating 32-column chunks
constexpr size_t SMEM2_BYTES = (size_t) STAGES2 * STAGE2_BYTES + 1024 + 256 + 2 * EPI_BYTES;

__device__ __forceinline__ uint32_t cluster_rank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank)      // shared::cluster address of `smem_addr` in CTA `rank`
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void mma_tf32_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                 :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar)           // arrives on `bar` in BOTH CTAs of the pair
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(smem_u32(bar)), "h"((uint16_t) 3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr)
{
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(bar_cluster_addr) : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc_m(int bm, int bn)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t) (bn >> 3) << 17) | ((uint32_t) (bm >> 4) << 24);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS2, 1)
gemm_tf32_2sm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                     float* __restrict__ C, long long ldc, int M, int N, int K,
                     const float* __restrict__ addend, long long ld_add, float alpha)
{
    constexpr int BN = BN2, STAGES = STAGES2;
    constexpr uint32_t TMEM_COLS = ACC_STAGES * BN;                           // all 512 columns
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t) 1023);
    unsigned char* tiles = smem;                                              // [STAGES][A half 16 KB | B half 16 KB]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t) STAGES * STAGE2_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* acc_full = empty + STAGES;
    uint64_t* acc_empty = acc_full + ACC_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + ACC_STAGES);
    float* epi = reinterpret_cast<float*>(smem + (size_t) STAGES * STAGE2_BYTES + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    const bool leader = rank == 0;
    const int pair = (int) (blockIdx.x >> 1), npairs = (int) (gridDim.x >> 1);
    const int tiles_m = (M + 2 * BM - 1) / (2 * BM), tiles_n = (N + BN - 1) / BN;
    const int ntiles = tiles_m * tiles_n;
    const int kblocks = (K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < ACC_STAGES; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 2 * EPI_WARPS2); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();                                                       // barriers and tensor memory of BOTH CTAs are ready
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ================= TMA producer (both CTAs; the transaction bytes land on the leader's barrier) =================
            int s = 0; uint32_t ph = 0;
            for (int t = pair; t < ntiles; t += npairs) {
                const int m0 = (t / tiles_n) * (2 * BM) + (int) rank * BM;
                const int n0 = (t % tiles_n) * BN + (int) rank * (BN / 2);
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&empty[s], ph ^ 1u);
                    if (leader) mbar_expect_tx(&full[s], 2 * STAGE2_BYTES);
                    const uint32_t bar = map_to_cta(smem_u32(&full[s]), 0);
                    unsigned char* a_dst = tiles + (size_t) s * STAGE2_BYTES;
                    tma_load_2d_2sm(a_dst, &map_a, bar, kb * BK, m0);
                    tma_load_2d_2sm(a_dst + A_BYTES, &map_b, bar, kb * BK, n0);
                    if (++s == STAGES) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {
            // ================= MMA issuer (leader CTA only) =================
            constexpr uint32_t idesc = make_idesc_m(2 * BM, BN);
            int s = 0; uint32_t ph = 0;
            int as = 0; uint32_t aph = 0;
            for (int t = pair; t < ntiles; t += npairs) {
                mbar_wait(&acc_empty[as], aph ^ 1u);                 // both CTAs' epilogues have drained this accumulator stage
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t) (as * BN);
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(tiles + (size_t) s * STAGE2_BYTES);
                    const uint32_t b_addr = a_addr + A_BYTES;
                    #pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t da = make_desc(a_addr + (uint32_t) (k * UMMA_K * 4));
                        const uint64_t db = make_desc(b_addr + (uint32_t) (k * UMMA_K * 4));
                        mma_tf32_2sm(tmem_d, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit_2sm(&empty[s]);                      // the stage is free in both CTAs once these MMAs have read it
                    if (kb == kblocks - 1) umma_commit_2sm(&acc_full[as]);
                    if (++s == STAGES) { s = 0; ph ^= 1u; }
                }
                if (++as == ACC_STAGES) { as = 0; aph ^= 1u; }
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue (each CTA drains its own 128 rows x 256 columns) =================
        const int q = warp & 3;                                      // TMEM lane quadrant this warp may read
        const int half = (warp - 4) >> 2;                            // which of the two warps of the quadrant
        int as = 0; uint32_t aph = 0;
        const uint32_t acc_empty_leader0 = map_to_cta(smem_u32(&acc_empty[0]), 0);
        for (int t = pair; t < ntiles; t += npairs) {
            const int m0 = (t / tiles_n) * (2 * BM) + (int) rank * BM, n0 = (t % tiles_n) * BN;
            mbar_wait(&acc_full[as], aph);
            tc_fence_after();
            float* stage = epi + (size_t) (warp - 4) * 32 * EPI_LD;
            #pragma unroll 1
            for (int c = half * 32; c < BN; c += 64) {
                uint32_t v[32];
                float ad[32];
                const int col = n0 + c + lane;
                if (addend) {
                    #pragma unroll
                    for (int r = 0; r < 32; ++r) {
                        const int row = m0 + q * 32 + r;
                        ad[r] = (row < M && col < N) ? __ldg(addend + (long long) row * ld_add + col) : 0.f;
                    }
                }
                const uint32_t taddr = tmem_base + ((uint32_t) (q * 32) << 16) + (uint32_t) (as * BN + c);
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                             "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                             "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                               "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                               "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                               "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                             : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                #pragma unroll
                for (int j = 0; j < 32; ++j) stage[lane * EPI_LD + j] = __uint_as_float(v[j]);
                __syncwarp();
                if (addend) {
                    #pragma unroll
                    for (int r = 0; r < 32; ++r) {
                        const int row = m0 + q * 32 + r;
                        if (row < M && col < N) C[(long long) row * ldc + col] = fmaf(alpha, ad[r], stage[r * EPI_LD + lane]);
                    }
                } else {
                    #pragma unroll 4
                    for (int r = 0; r < 32; ++r) {
                        const int row = m0 + q * 32 + r;
                        if (row < M && col < N) C[(long long) row * ldc + col] = stage[r * EPI_LD + lane];
                    }
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(acc_empty_leader0 + (uint32_t) (as * sizeof(uint64_t)));
            if (++as == ACC_STAGES) { as = 0; aph ^= 1u; }
        }
    }
    tc_fence_before();
    cluster_sync_all();                                    // nobody leaves (or frees tensor memory) while the pair still works
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---- host side ------------------------------------------------------------------------------------------------
struct Host {
    PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
    int sms = 0;
    int max_pairs = 0;            // CTA pairs (clusters of 2) of the 2-SM kernel that can be resident at once; 0: kernel not usable
    bool ok = false;
    bool tried = false;
};

inline Host& host()
{
    static Host h;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (h.tried) return h;
    h.tried = true;
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return h;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&h.sms, cudaDevAttrMultiProcessorCount, dev);
    if (major != 10) return h;                                    // tcgen05 / TMEM: sm_100 family only
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) { cudaGetLastError(); return h; }
    h.encode = (PFN_cuTensorMapEncodeTiled_v12000) fn;
    if (cudaFuncSetAttribute(gemm_tf32_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) Cfg<128>::SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(gemm_tf32_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) Cfg<256>::SMEM_BYTES) != cudaSuccess) { cudaGetLastError(); return h; }
    h.ok = true;
    if (cudaFuncSetAttribute(gemm_tf32_2sm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) SMEM2_BYTES) == cudaSuccess) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned) (h.sms & ~1)); cfg.blockDim = dim3(THREADS2); cfg.dynamicSmemBytes = SMEM2_BYTES;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int nc = 0;
        if (cudaOccupancyMaxActiveClusters(&nc, gemm_tf32_2sm_kernel, &cfg) == cudaSuccess && nc > 0) h.max_pairs = nc;
    }
    cudaGetLastError();
    return h;
}

inline bool make_map(Host& h, CUtensorMap* map, const float* base, long long ld, int rows, int cols, int box_rows)
{
    cuuint64_t dims[2] = {(cuuint64_t) cols, (cuuint64_t) rows};
    cuuint64_t strides[1] = {(cuuint64_t) ld * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t) BK, (cuuint32_t) box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = h.encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

}  // namespace tf32gemm

// worth it and legal for TMA? (16-byte aligned bases, leading dimensions multiples of 4 floats, a few tiles of work)
inline bool sm100_gemm_tf32_usable(const float* A, long long lda, const float* B, long long ldb, const float* C, long long ldc,
                                   int M, int N, int K)
{
    (void) C; (void) ldc;
    if ((((uintptr_t) A) | ((uintptr_t) B)) & 15u) return false;
    if ((lda & 3) || (ldb & 3)) return false;
    if (M < 64 || N < 64 || K < 64) return false;
    return (double) M * (double) N * (double) K >= 64.0 * 1024 * 1024;
}

// returns 0 = launched, 1 = tensor-core path not available here (caller falls back), < 0 = error
inline int sm100_gemm_tf32(const float* A, long long lda, const float* B, long long ldb, float* C, long long ldc,
                           int M, int N, int K, cudaStream_t st, const float* addend = nullptr, long long ld_add = 0,
                           float alpha = 0.f, const tf32gemm::ScatterArgs* scatter = nullptr)
{
    using namespace tf32gemm;
    Host& h = host();
    if (!h.ok) return 1;
    static int force_bn = -1;
    if (force_bn < 0) { const char* e = getenv("STOCHQN_B200_GEMM_BN"); force_bn = e ? atoi(e) : 0; }      // dev switch
    static int use_2sm = -1;
    if (use_2sm < 0) { const char* e = getenv("STOCHQN_B200_GEMM_2SM"); use_2sm = e ? atoi(e) : 1; }                 // dev switch
    const bool no_scatter = !scatter || scatter->world <= 1;
    if (use_2sm && h.max_pairs > 0 && no_scatter && force_bn == 0 && M >= 256 && N >= 256) {
        // CTA pairs on 256 x 256 tiles whenever they give most pairs work
        const long long nt = (long long) ((M + 2 * BM - 1) / (2 * BM)) * ((N + BN2 - 1) / BN2);
        if (nt * 4 >= (long long) h.max_pairs * 3 || use_2sm == 2) {
            CUtensorMap ma2, mb2;
            if (make_map(h, &ma2, A, lda, M, K, BM) && make_map(h, &mb2, B, ldb, N, K, BN2 / 2)) {
                const int pairs = (int) (nt < h.max_pairs ? nt : h.max_pairs);
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3((unsigned) (2 * pairs)); cfg.blockDim = dim3(THREADS2); cfg.dynamicSmemBytes = SMEM2_BYTES; cfg.stream = st;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                if (cudaLaunchKernelEx(&cfg, gemm_tf32_2sm_kernel, ma2, mb2, C, ldc, M, N, K, addend, ld_add, alpha) == cudaSuccess) return 0;
                cudaGetLastError();             // fall through to the one-CTA kernel
            }
        }
    }
    const long long tiles_m = (M + BM - 1) / BM;
    // wide tiles when they still give every SM work (the wide kernel makes the same makespan with less L2 traffic)
    const bool wide = force_bn == 256 || (force_bn != 128 && N >= 256 && tiles_m * ((N + 255) / 256) >= (long long) h.sms * 3 / 4);
    CUtensorMap ma, mb;
    if (!make_map(h, &ma, A, lda, M, K, BM) || !make_map(h, &mb, B, ldb, N, K, wide ? 256 : 128)) return 1;
    const long long ntiles = tiles_m * ((N + (wide ? 255 : 127)) / (wide ? 256 : 128));
    const int grid = (int) (ntiles < h.sms ? ntiles : h.sms);
    const ScatterArgs sc = scatter ? *scatter : ScatterArgs();
    if (wide) gemm_tf32_kernel<256><<<grid, THREADS, Cfg<256>::SMEM_BYTES, st>>>(ma, mb, C, ldc, M, N, K, addend, ld_add, alpha, sc);
    else      gemm_tf32_kernel<128><<<grid, THREADS, Cfg<128>::SMEM_BYTES, st>>>(ma, mb, C, ldc, M, N, K, addend, ld_add, alpha, sc);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
