// internal_rs.h - shared between multinomial.cu (the fused GEMM -> reduce-scatter) and ext_impl.inc (the communicator).
#pragma once
#include <cuda_runtime.h>

struct StochqnRsPlan {
    int world, rank;
    long long blk;                 // elements per owner block
    void* dst[16];                 // dst[o]: rank o's receive slot for THIS sender in the current parity (peer-mapped)
    void* local;                   // this rank's `world` receive slots of the current parity, blk elements each
};
// start a fused reduce-scatter (collective); -5 when the communicator has no peer-memory path
int stochqn_b200_internal_rs_begin(void* comm, long long blk, StochqnRsPlan* out);
// rank barrier on `st` (peer-memory flags; ncclAllReduce of one value as fallback)
int stochqn_b200_internal_barrier(void* comm, cudaStream_t st);

// multinomial.cu: where the one-launch small-batch gradient (mn_small.cuh) keeps its scratch inside a caller's `work` buffer for a
// batch of B rows, and whether the shape qualifies on a grid of `grid` CTAs (ok = 1); smem = dynamic shared memory it needs
struct StochqnMnSmallPlan {
    size_t off_zp, off_dg, smem;
    int ok;
};
void stochqn_b200_internal_mn_small_plan(long long B, long long d, long long K, int grid, StochqnMnSmallPlan* out);
