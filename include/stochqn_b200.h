/*  stochqn_b200.h - device-side extensions of the stochQN C ABI (B200 / sm_100a build)
 *
 *  stochqn.h is the drop-in part: the reference's nine entry points with unchanged
 *  signatures.  This header adds what a GPU-resident caller needs and the reference,
 *  being a host library, never had: device / stream selection, sharding of one
 *  optimizer over several GPUs, the bundled device callbacks (Rosenbrock; binary
 *  logistic gradient / Hessian-vector / loss of the reference's R/logistic.R:1-37), and
 *  workspace export / import in the reference's own field layout.
 *
 *  Plain C: pointers and sizes only, no CUDA or torch types in any signature
 *  (`void *stream` is a cudaStream_t, `void *comm` an opaque handle).
 *  Every function returns 0 on success and a negative code on failure unless stated;
 *  stochqn_b200_last_error() gives the message.  `ws` is a pointer returned by
 *  initialize_oLBFGS / initialize_SQN / initialize_adaQN.
 */
#ifndef STOCHQN_B200_INCLUDE
#define STOCHQN_B200_INCLUDE

#include "stochqn.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library information -------------------------------------------------------------- */
int         stochqn_b200_version(void);          /* 100*major + minor */
int         stochqn_b200_real_bytes(void);       /* sizeof(real_t) of this build: 8 or 4 */
const char* stochqn_b200_last_error(void);       /* message of the last failure on this thread */
/* number of CUDA kernels this library has launched since it was loaded (all workspaces) */
unsigned long long stochqn_b200_launch_count(void);

/* ---- device / stream ------------------------------------------------------------------- */
/* initialize_*() allocates on the CUDA device that is current when it is called.  All work
   of a workspace is enqueued on one stream: the legacy default stream unless set here
   (pass e.g. torch.cuda.current_stream().cuda_stream).  With host pointers run_*() returns after
   that stream has drained, so results are final on return, as with the reference; with device
   pointers results are final in stream order (see STOCHQN_B200_OPT_SYNC_RETURN). */
int stochqn_b200_set_stream(void *ws, void *stream);

/* ---- per-workspace options --------------------------------------------------------------- */
enum stochqn_b200_option {
    /* 1: `grad` holds the search direction after a step (oLBFGS: -step*direction), exactly what
       the reference leaves there (src/stochqn.c:838,1006).  0: `grad` is left untouched - saves
       one n-vector write (device pointers) or one device->host copy of n values (host pointers)
       per step; the reference documents the array only as an input that "will be modified in
       place" (include/stochqn.h:342, 356-358), not as an output.  -1 (default): automatic -
       1 for device-pointer calls (the write is nearly free), 0 for host-pointer calls. */
    STOCHQN_B200_OPT_GRAD_WRITEBACK = 1,
    /* host-pointer calls only.  1 (default): the device mirror of `x` is trusted between calls -
       `x` is uploaded once (and again whenever another array is passed, or after the library
       itself rewrote it), then only downloaded after each step.  This is the reference's own
       contract: *req == x "do NOT modify the values in this array" (include/stochqn.h:364-366;
       stochqn/_optimizers.py:995-999).  0: upload `x` on every call that reads it, for callers
       that do modify x between calls (projections, restarts). */
    STOCHQN_B200_OPT_TRUST_X_MIRROR = 2,
    /* 1: bracket the streaming kernels of each call (K1 multi-dot, K3 combine/update, K4 pair; adaQN: KA1, KA3, KA2) with CUDA
       events on the workspace stream and accumulate their device times; setting it (to 0 or 1) resets the
       accumulators.  Read them with stochqn_b200_get_stat.  Default 0. */
    STOCHQN_B200_OPT_PROFILE = 3,
    /* device-pointer calls only.  0 (default): run_*() returns as soon as the flags its control flow needs have
       come back (accept / reject of the direction, the curvature dots); the streaming kernel that updates x may
       still be running, and the results are final in the order of the workspace stream - enqueue the next
       gradient evaluation on that stream (or synchronise it) as with any CUDA library.  1: drain the stream
       before every return, the reference's contract.  Host-pointer calls always return with everything final. */
    STOCHQN_B200_OPT_SYNC_RETURN = 4,
    /* oLBFGS / SQN: largest n for which a step is ONE launch (dots, solve and update fused - for the latency-bound
       sizes where three launches dominate) instead of K1 -> K2 -> K3: a single 1024-thread CTA up to n = 2048, a
       cooperative grid with one grid barrier above.  Default 2048 (environment: STOCHQN_B200_SMALL_N); 0 disables.
       Not used when the optimizer is sharded. */
    STOCHQN_B200_OPT_ONE_LAUNCH_MAX_N = 5,
    /* stochqn_b200_fit_batch / fit_batches: largest n for which the request loop of a mini-batch runs on the device
       (csrc/kernels_loop.cuh; default 65536 for oLBFGS / SQN and 524288 for adaQN, whose ordinary steps then are ONE launch -
       kl_ada - for mem_size <= 20; environment: STOCHQN_B200_LOOP_MAX_N); 0: always the host-driven loop. */
    STOCHQN_B200_OPT_DEVICE_LOOP_MAX_N = 6,
    /* stochqn_b200_fit_batch / fit_batches, two-class models (0, 1), oLBFGS and the ordinary steps of SQN, n <= 5120:
       1 (default; environment STOCHQN_B200_FUSED_FIT): a whole RUN of consecutive mini-batches is ONE cooperative launch
       (csrc/kernels_fit.cuh) - the gradient sweeps, the step, the pair and every decision between them, separated by grid
       barriers instead of launches; 0: one kernel sequence per mini-batch (csrc/kernels_loop.cuh). */
    STOCHQN_B200_OPT_FUSED_FIT = 7
};
int stochqn_b200_set_option(void *ws, int option, long long value);
/* current value of an option of this workspace (-1: unknown option or workspace) */
long long stochqn_b200_get_option(void *ws, int option);
/* development aid (tools/probe_fit.py): a device buffer of 16 * nbatches 64-bit slots in which CTA 0 of the fused fit kernel
   (STOCHQN_B200_OPT_FUSED_FIT) leaves %globaltimer stamps at the phase boundaries of every mini-batch; NULL turns it off */
int stochqn_b200_debug_fit_trace(void *ws, unsigned long long *dev_buf);
/* the same for the one-launch small-batch multinomial gradient (8 slots, process-wide) and - through stochqn_b200_debug_fit_trace -
   for the one-launch adaQN step (16 slots per call, overwritten by every call) */
int stochqn_b200_debug_mn_trace(unsigned long long *dev_buf);

enum stochqn_b200_stat {
    STOCHQN_B200_STAT_K1_MS = 1, STOCHQN_B200_STAT_K1_COUNT = 2,     /* accumulated device ms / launches (profile mode) */
    STOCHQN_B200_STAT_K3_MS = 3, STOCHQN_B200_STAT_K3_COUNT = 4,
    STOCHQN_B200_STAT_K4_MS = 5, STOCHQN_B200_STAT_K4_COUNT = 6,
    STOCHQN_B200_STAT_LAST_BOUND = 7,                                 /* bound on ||direction|| of the last step */
    STOCHQN_B200_STAT_EXACT_NORM_STEPS = 8,                           /* steps that needed the exact-norm (two-pass) route */
    STOCHQN_B200_STAT_KA2_MS = 9, STOCHQN_B200_STAT_KA2_COUNT = 10,   /* adaQN: the second dot pass (K1/K3 slots hold KA1/KA3) */
    /* steps taken by the one-launch kernel used for latency-bound sizes (n <= 2048 unless STOCHQN_B200_SMALL_N
       says otherwise; 0 there disables it): dots, solve and update in one cooperative launch instead of three */
    STOCHQN_B200_STAT_ONE_LAUNCH_STEPS = 11,
    STOCHQN_B200_STAT_DEVICE_LOOP_STEPS = 12,                         /* steps taken by the device-side loop kernels */
    STOCHQN_B200_STAT_FUSED_FIT_STEPS = 13                            /* ... of which inside fused runs (kernels_fit.cuh) */
};
int stochqn_b200_get_stat(void *ws, int what, double *out);

/* leading dimension (in elements) of s_mem / y_mem / F rows: rows are padded to a multiple
   of 128 bytes so that every row is 16-byte aligned whatever n is; row j of s_mem starts
   at s_mem + j * row_stride. */
size_t stochqn_b200_row_stride(void *ws);

/* ---- sharding one optimizer over several GPUs (one process per GPU) ---------------------
   Every n-vector is split into contiguous blocks; each rank creates its workspace with
   n = its own block length and registers a communicator.  Inside a step the only exchange
   is one small all-reduce (sum, fp64, <= 4*mem_size+2 values) per reduction phase, fused into
   the kernel that sums the partial dots (peer-memory mailboxes over NVLink; the ranks must call
   the library in the same order, and every rank must own its GPU).
   Bootstrap: rank 0 obtains a 128-byte id, the host program ships it to the other ranks
   (torch.distributed broadcast, MPI, a file ...), every rank calls comm_init. */
int stochqn_b200_comm_unique_id(void *id128);
int stochqn_b200_comm_init(const void *id128, int rank, int world_size, void **comm);
/* comm_destroy frees this rank's mailbox and peer-mapped vectors: let every rank finish its device work and meet at a
   host barrier first (a peer's kernel may still be reading them) */
int stochqn_b200_comm_destroy(void *comm);
/* 1 when the small all-reduces of this communicator are done over NVLink peer memory inside the kernel that
   produces the values (cudaIpc mailboxes, one rank per GPU), 0 when they go through ncclAllReduce
   (no peer access between the devices, or STOCHQN_B200_NO_P2P=1 in the environment) */
int stochqn_b200_comm_uses_p2p(void *comm);
/* n_global = sum of all ranks' n: the reference's step-rejection limit is 1e3 * n
   (src/stochqn.c:829) and must use the length of the whole vector */
int stochqn_b200_set_comm(void *ws, void *comm, long long n_global);
/* in-place sum of `count` doubles in device memory over the communicator, on `stream`
   (exposed so that callbacks can ride on the same communicator) */
int stochqn_b200_allreduce_f64(void *comm, double *dev_buf, size_t count, void *stream);

/* n-vector collectives for row-sharded gradient / Hessian-vector callbacks (each rank evaluates the callback on its rows
   of the batch; NCCL over NVLink, on `stream`).  all-reduce: every rank ends with the full sum (replicated optimizer).
   reduce-scatter + all-gather: rank r receives block r (block_count elements; the vector is world_size * block_count
   long) of the sum, steps its shard of the optimizer state (stochqn_b200_set_comm) and the updated x is gathered back. */
int stochqn_b200_allreduce_real(void *comm, real_t *dev_buf, size_t count, void *stream);
int stochqn_b200_reduce_scatter_real(void *comm, const real_t *send_full, real_t *recv_block, size_t block_count, void *stream);
int stochqn_b200_all_gather_real(void *comm, const real_t *send_block, real_t *recv_full, size_t block_count, void *stream);

/* all-gather over NVLink peer memory: ONE kernel that stores this rank's block into every rank's copy of the gathered
   vector, then the rank barrier.  *gathered (world_size * block_count elements) belongs to the library and stays valid,
   in stream order, until the second next call on this communicator (double-buffered).  -5: no peer-memory path - use
   stochqn_b200_all_gather_real. */
int stochqn_b200_all_gather_p2p(void *comm, const real_t *send_block, size_t block_count, real_t **gathered, void *stream);

/* reduce-scatter over NVLink peer memory, by pulling: the rank barrier (every rank's vector is complete), then ONE kernel
   in which the owner of block r reads block r of every rank's send vector over NVLink (16-byte loads, all world_size of
   them in flight) and adds them in rank order (deterministic).  The send vector has to be mapped by the peers, so it
   belongs to the library: stochqn_b200_p2p_send_buffer returns the vector (world_size * block_count elements) that the
   NEXT reduce_scatter_p2p call on this communicator reads - let the producing kernel (e.g. multinomial_loss_grad) write
   straight into it; any other send_full is copied into it first (one extra pass).  Double-buffered: the pointer changes
   from call to call, ask again after every reduce_scatter_p2p.  Replaces stochqn_b200_reduce_scatter_real
   (ncclReduceScatter).  -5: no peer-memory path.  Collective: the first call (and a call with another block_count)
   allocates on every rank. */
int stochqn_b200_p2p_send_buffer(void *comm, size_t block_count, real_t **send_full);
int stochqn_b200_reduce_scatter_p2p(void *comm, const real_t *send_full, real_t *recv_block, size_t block_count, void *stream);

/* world_size communicators whose ranks all live in THIS process on the current device (comms[0..world_size)): the
   peer-memory collectives above, the mailbox all-reduce and the fused reduce-scatter run the very kernels of the
   one-process-per-GPU case, with the peers' buffers shared by pointer instead of cudaIpc - how the exchange kernels are
   tested on a single-GPU box.  Give every rank its own non-blocking stream and do not synchronise the host between the
   ranks' calls of one collective (the waiting kernels of all ranks must be running together).  Two more conditions of
   sharing one device: run with CUDA_MODULE_LOADING=EAGER (the first launch of a lazily loaded kernel may wait for the
   device to drain, which a polling kernel prevents) and CUDA_DEVICE_MAX_CONNECTIONS >= the number of rank streams (on a
   shared hardware queue rank 1's launch sits behind the kernel that follows rank 0's polling kernel).  The first use of
   a collective allocates for the whole group and drains the device: issue it while nothing is waiting.  There is no NCCL behind
   such a communicator: the *_real collectives answer -5.  Each one is released with stochqn_b200_comm_destroy. */
int stochqn_b200_comm_init_inprocess(int world_size, void **comms);
/* 1 once a stand-alone exchange on this communicator gave up waiting for a peer (20 s) - its result is then garbage */
int stochqn_b200_comm_error(void *comm);

/* ---- bundled device callbacks -------------------------------------------------------------
   Chained Rosenbrock, formulas of the reference's example (example/c_rosen.c:13-41), on a
   contiguous shard x[0..n_local) of a vector of length n_global that starts at global index
   `offset`.  `halo` = {x[offset-1], x[offset+n_local]} in DEVICE memory (ignored at the ends
   of the global vector; may be NULL when the shard is the whole vector). */
int stochqn_b200_rosenbrock_x0(real_t *x, long long n_local, long long offset, void *stream);
int stochqn_b200_rosenbrock_grad(const real_t *x, real_t *grad, long long n_local, long long offset,
                                 long long n_global, const real_t *halo, void *stream);
/* writes the local part of the objective (fp64) to *f_dev (device memory) */
int stochqn_b200_rosenbrock_fun(const real_t *x, long long n_local, long long offset, long long n_global,
                                const real_t *halo, double *f_dev, void *stream);

/* Sharded runs: fills halo[0] = last element of the left neighbour's shard and halo[1] = first element
   of the right neighbour's (device memory, real_t[2]) with ONE small all-reduce on `stream`;
   `scratch` is device memory for 2*world_size doubles. */
int stochqn_b200_rosenbrock_halo(const real_t *x, long long n_local, int rank, int world_size, void *comm,
                                 real_t *halo, double *scratch, void *stream);

/* Binary logistic regression, the closed forms of R/logistic.R:1-37.  X is ROW-major
   [nrows][ncols] with leading dimension ldx (a batch is a row range: no copy); y in {0,1};
   sample weights `sw` may be NULL.  All pointers are device pointers.
     grad     = X'((sigmoid(Xw) - y) .* sw) / sum(sw) + 2*lambda*w          (logistic.R:12-21)
     hess_vec = X'( p(1-p) .* sw .* (Xv) ) / sum(sw) + 2*lambda*v           (logistic.R:23-37)
     loss     = -sum(sw .* (y log p + (1-y) log(1-p))) / sum(sw) + lambda*|w|^2   (logistic.R:1-10)
   `work` is device scratch of at least stochqn_b200_logistic_work_size(nrows, ncols) bytes. */
size_t stochqn_b200_logistic_work_size(long long nrows, long long ncols);

/* Sparse model matrices (stochqn/_logistic.py:155 keeps scipy CSR inputs): rows [row0, row0 + nrows) of a device-resident
   canonical CSR matrix (indptr / indices as 64-bit integers, no duplicate column within a row) expanded into dense rows
   out[r][0..ncols), leading dimension ldo - the layout the bundled callbacks stream.  *bad_index_flag (device int, may be
   NULL) is set to 1 if a column index lies outside [0, ncols). */
int stochqn_b200_csr_to_dense(const long long *indptr, const long long *indices, const real_t *data, long long row0, long long nrows,
                              long long ncols, real_t *out, long long ldo, int *bad_index_flag, void *stream);
int stochqn_b200_logistic_grad(const real_t *X, long long ldx, const real_t *y, const real_t *sw,
                               long long nrows, long long ncols, const real_t *w, real_t lambda,
                               real_t *grad, void *work, void *stream);
int stochqn_b200_logistic_hess_vec(const real_t *X, long long ldx, const real_t *y, const real_t *sw,
                                   long long nrows, long long ncols, const real_t *w, const real_t *v,
                                   real_t lambda, real_t *hess_vec, void *work, void *stream);
int stochqn_b200_logistic_loss(const real_t *X, long long ldx, const real_t *y, const real_t *sw,
                               long long nrows, long long ncols, const real_t *w, real_t lambda,
                               double *loss_dev, void *work, void *stream);

/* The same three closed forms in the conventions of the scikit-learn (<= 1.0) private functions the reference's Python
   layer calls for two-class problems (stochqn/_logistic.py:23-30: _logistic_loss_and_grad, _logistic_grad_hess):
   labels y in {-1,+1}; SUMS over samples (the caller pre-normalises sw, stochqn/_logistic.py:167); w has
   ncols + fit_intercept entries, the intercept LAST and unpenalised; penalty alpha/2 * |w[:ncols]|^2.  With
   z_i = x_i'w[:ncols] + intercept and q_i = sigmoid(y_i z_i):
     loss     = -sum_i sw_i log q_i + alpha/2 |w[:ncols]|^2
     grad     = [ X'r + alpha w[:ncols] ; sum(r) ]            r_i = sw_i (q_i - 1) y_i
     hess_vec = [ X'r + alpha v[:ncols] ; sum(r) ]            r_i = sw_i q_i (1 - q_i) (x_i'v[:ncols] + v[ncols])
   Same kernels, same `work` size as above (one sweep of the batch for ncols <= 5120). */
int stochqn_b200_logistic_sk_grad(const real_t *X, long long ldx, const real_t *y, const real_t *sw,
                                  long long nrows, long long ncols, int fit_intercept, const real_t *w,
                                  real_t alpha, real_t *grad, void *work, void *stream);
int stochqn_b200_logistic_sk_hess_vec(const real_t *X, long long ldx, const real_t *y, const real_t *sw,
                                      long long nrows, long long ncols, int fit_intercept, const real_t *w,
                                      const real_t *v, real_t alpha, real_t *hess_vec, void *work, void *stream);
int stochqn_b200_logistic_sk_loss(const real_t *X, long long ldx, const real_t *y, const real_t *sw,
                                  long long nrows, long long ncols, int fit_intercept, const real_t *w,
                                  real_t alpha, double *loss_dev, void *work, void *stream);

/* Sharded gradient in ONE launch: the halo exchange is fused into the gradient kernel (CTA 0 exchanges the shard
   ends over the communicator's peer-memory mailboxes while the other CTAs stream; falls back to
   stochqn_b200_rosenbrock_halo + stochqn_b200_rosenbrock_grad when the communicator has no peer-memory path).
   `halo` (real_t[2]) and `scratch` (2*world_size doubles) are device scratch for the fallback. */
int stochqn_b200_rosenbrock_grad_sharded(const real_t *x, real_t *grad, long long n_local, long long offset,
                                         long long n_global, int rank, int world_size, void *comm,
                                         real_t *halo, double *scratch, void *stream);

/* Multinomial (softmax) logistic regression with the semantics of the scikit-learn (<= 1.0) private functions the
   reference's Python layer calls (stochqn/_logistic.py:7-13: _multinomial_loss_grad, _multinomial_grad_hess):
   w is (nclasses x (nfeat + fit_intercept)) row-major, intercept = last column; X row-major [nrows][nfeat] with leading
   dimension ldx; the targets are either a dense indicator / probability matrix Y [nrows][nclasses] (leading dimension
   ldy) or, with Y == NULL, int32 class labels; sample weights `sw` may be NULL.  SUMS over samples (not means):
     loss     = -sum_i sw_i sum_k Y_ik log softmax(X w' + b)_ik + alpha/2 * ||w[:, :nfeat]||^2
     grad     = (sw .* (P - Y))' X + alpha w ;  intercept column = column sums of sw .* (P - Y)
     hess_vec = R' X + alpha v with R = sw .* P .* (X v' + vb - rowsum(P .* (X v' + vb))) ; intercept: column sums of R
   `grad` or `loss_dev` (device double) may be NULL.  `work`: device scratch of stochqn_b200_multinomial_work_size bytes.
   The two matrix products run on the tensor cores (tcgen05, tf32 inputs, fp32 accumulation: relative error <= 2^-10
   per product) in the float build when they are large and 16-byte aligned; STOCHQN_B200_NO_TENSOR_CORES=1 in the
   environment, or the double build, keeps them on the CUDA cores in full precision. */
size_t stochqn_b200_multinomial_work_size(long long nrows, long long nfeat, long long nclasses);
/* the product both callbacks are built on, exposed for tests and probes: C[M x N] (ldc) = A[M x K] (lda) * B[N x K]' (ldb),
   row-major, device pointers; same tensor-core / CUDA-core selection as above */
int stochqn_b200_gemm_tn(const real_t *A, long long lda, const real_t *B, long long ldb, real_t *C, long long ldc,
                         int M, int N, int K, void *stream);
int stochqn_b200_multinomial_loss_grad(const real_t *X, long long ldx, const real_t *Y, long long ldy, const int *labels,
                                       const real_t *sw, long long nrows, long long nfeat, long long nclasses,
                                       int fit_intercept, const real_t *w, real_t alpha, real_t *grad, double *loss_dev,
                                       void *work, void *stream);
int stochqn_b200_multinomial_hess_vec(const real_t *X, long long ldx, const real_t *Y, long long ldy, const int *labels,
                                      const real_t *sw, long long nrows, long long nfeat, long long nclasses,
                                      int fit_intercept, const real_t *w, const real_t *v, real_t alpha, real_t *hess_vec,
                                      void *work, void *stream);

/* Row-sharded gradient with the reduce-scatter FUSED into the product that computes it (float build, peer-memory
   communicator): every rank evaluates the multinomial gradient on ITS rows of the batch (pass alpha / world_size, the
   penalty is then added once in the sum); the epilogue of the second GEMM does not store the gradient locally but
   writes each tile straight into the receive slot of the rank that owns that block of the n-vector
   (n = world_size * block_count, block r = elements [r*block_count, (r+1)*block_count)) over NVLink, while the tensor
   pipe works on the next tile; a rank barrier and one pass that adds the world_size slots (rank order, deterministic)
   leave block `rank` of the summed gradient in grad_block.  Replaces multinomial_loss_grad + reduce_scatter_real.
   Returns -5 when this path is not available (double build, no peer access, shapes the tensor-core kernel does not
   take): call the two-step form then.  Collective: every rank must call it in the same order. */
int stochqn_b200_multinomial_grad_reduce_scatter(void *comm, const real_t *X, long long ldx, const real_t *Y, long long ldy,
                                                 const int *labels, const real_t *sw, long long nrows, long long nfeat,
                                                 long long nclasses, int fit_intercept, const real_t *w, real_t alpha,
                                                 real_t *grad_block, long long block_count, void *work, void *stream);

/* ---- guided mode: the request loop of one mini-batch, natively ------------------------------------------------
   The reference's guided classes serve the optimizer's requests from the host language: one interpreter round trip
   per request (R/optimizers_guided.R:26-111 `run_stochQN_on_batch`; stochqn/_optimizers.py:339-382 `_fit_batch`).
   stochqn_b200_fit_batch runs that loop inside the library for a DEVICE-resident model matrix and the bundled
   callbacks: starting from the pending request (*task, *req, *req_vec - as left by the previous run_*() or
   fit_batch call), it evaluates what is asked on the right rows, hands it to run_oLBFGS / run_SQN / run_adaQN, and
   repeats until the optimizer asks for a gradient on a NEW batch (task calc_grad), exactly the reference's loop:
     calc_grad, calc_grad_same_batch      gradient on `batch`
     calc_grad_big_batch, calc_hess_vec   gradient / Hessian-vector product on `long_batch`
     calc_fun_val_batch                   objective on `valset` if given, else on `long_batch`
   Every batch is a ROW RANGE of resident arrays (pointer + row count: no copy; the "long batch" of the last
   bfgs_upd_freq mini-batches is one range when they are consecutive rows).  All pointers are device pointers.
   Returns 0 when the loop ended with *task == calc_grad; 1 when a request needs `long_batch` (or `valset`) and
   none was given, or `long_batch` was already used once in this call (the reference's stash is emptied by the
   first use, stochqn/_optimizers.py:92-107) - *task / *req / *req_vec then describe the pending request and the
   caller serves it its own way; negative on failure (-1000: invalid workspace, as run_*()). */
typedef struct {
    const real_t *X;        /* [nrows][ncols] row-major, leading dimension ldx */
    long long ldx;
    const real_t *y;        /* models 0, 1: labels [nrows]; model 2: one-hot matrix [nrows][nclasses], leading dimension ldy */
    long long ldy;
    const real_t *sw;       /* sample weights [nrows] or NULL */
    long long nrows;
} stochqn_b200_rows;

typedef struct {
    int model;              /* 0: two classes, R conventions (stochqn_b200_logistic_*);
                               1: two classes, scikit-learn conventions (stochqn_b200_logistic_sk_*);
                               2: multinomial (stochqn_b200_multinomial_*) */
    int fit_intercept;      /* models 1, 2 */
    long long ncols;        /* features (columns of X) */
    long long nclasses;     /* model 2 */
    real_t reg_param;
    void *work;             /* device scratch, at least the *_work_size of the largest batch handed over */
} stochqn_b200_model;

typedef struct {
    long long calls;        /* run_*() calls made */
    long long n_info[4];    /* how many of them reported no_problems / func_increased / curvature_too_small / search_direction_was_nan */
    int last_info;          /* info_enum of the last call */
    int x_changed;          /* 1 if any call updated x */
    int long_batch_used;    /* 1 if `long_batch` served a request */
} stochqn_b200_fit_report;

int stochqn_b200_fit_batch(void *ws, real_t *x, real_t step_size, const stochqn_b200_model *model,
                           const stochqn_b200_rows *batch, const stochqn_b200_rows *long_batch,
                           const stochqn_b200_rows *valset, int *task, real_t **req, real_t **req_vec,
                           stochqn_b200_fit_report *report);

/* Several consecutive mini-batches of one resident matrix in ONE call, with the request loop on the DEVICE:
   mini-batch b (b = 0 .. nbatches-1) is rows [first_row + b*batch_rows, first_row + (b+1)*batch_rows) of `data` (the
   last one may be cut short by data->nrows); the long batch a request of mini-batch b may need is rows
   [long_first[b], long_first[b] + long_rows[b]) of `data` (host arrays of nbatches entries, or NULL: none).
   For oLBFGS, and for the ordinary steps of SQN (those that are not followed by averaging / pair work), up to
   STOCHQN_B200_OPT_DEVICE_LOOP_MAX_N variables, nothing is waited for: the ring-buffer counters live in a device record,
   the step and pair kernels (csrc/kernels_loop.cuh) take the accept / reject / curvature decisions of
   src/stochqn.c:825-835, 883-900 themselves and skip what the reference would skip, and the host enqueues
   gradient -> step (-> gradient -> pair) per mini-batch back to back.  The counters of the public struct, the task /
   info tallies of `report` and *task / *req are brought up to date once, before the call returns.  Every other
   mini-batch (SQN / adaQN pair boundaries, adaQN, larger n, sharded workspaces) goes through the loop of
   stochqn_b200_fit_batch.  Same return values as stochqn_b200_fit_batch (1: a request of some mini-batch could not be
   served from the ranges given; the call stops there with that request pending).  stochqn_b200_fit_batch itself takes
   the same route for its single mini-batch (one wait per mini-batch instead of one per request). */
int stochqn_b200_fit_batches(void *ws, real_t *x, real_t step_size, const stochqn_b200_model *model,
                             const stochqn_b200_rows *data, long long first_row, long long batch_rows, long long nbatches,
                             const long long *long_first, const long long *long_rows, const stochqn_b200_rows *valset,
                             int *task, real_t **req, real_t **req_vec, stochqn_b200_fit_report *report);

/* ---- workspace export / import (checkpoint / resume) ----------------------------------------
   The reference keeps all state in host-language arrays, so saveRDS / pickle of the R / Python
   object was a checkpoint (R/allocators.R, stochqn/_optimizers.py:791-879).  These copy the same
   fields between the device workspace and dense host arrays in the reference's layout
   (s_mem / y_mem / F row-major [mem_size][n] without padding).  Pass NULL for fields that do not
   exist in the optimizer or are not wanted.  Scalars travel through the public struct itself. */
typedef struct {
    real_t *s_mem, *y_mem;          /* [mem_size][n] */
    real_t *grad_prev;              /* [n] */
    real_t *x_sum, *x_avg_prev;     /* [n]  (SQN, adaQN) */
    real_t *grad_sum_sq;            /* [n]  (adaQN) */
    real_t *F;                      /* [fisher_size][n] (adaQN, Fisher mode) */
} stochqn_b200_host_state;
int stochqn_b200_export(void *ws, stochqn_b200_host_state *out);
int stochqn_b200_import(void *ws, const stochqn_b200_host_state *in);

#ifdef __cplusplus
}
#endif
#endif /* STOCHQN_B200_INCLUDE */
