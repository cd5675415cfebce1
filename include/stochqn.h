/*  stochqn.h - public C ABI of the B200-native stochastic quasi-Newton step
 *
 *  Source-compatible replacement for the reference's public header
 *  (david-cortes/stochQN, include/stochqn.h).  Every struct keeps the reference's
 *  field names, order and types (reference include/stochqn.h:86-151), every enum
 *  keeps its numeric values (268-291) and the nine entry points keep their
 *  signatures (227-238, 381-383), so C, C++, R (.Call) and Cython callers compile
 *  against it unchanged.  What differs is where the data lives:
 *
 *    - every array pointer inside a workspace (s_mem, y_mem, grad_prev, x_sum,
 *      x_avg_prev, H0, grad_sum_sq, F ...) is a DEVICE pointer into HBM;
 *    - `x`, `grad`, `hess_vec` passed to run_*() may be device pointers (the
 *      native mode: nothing leaves the GPU except a few flag words per call) or
 *      ordinary host pointers (compatibility mode: the library stages them through
 *      pinned memory, and `*req` / `*req_vec` then point at host mirrors);
 *    - `buffer_rho`, `buffer_alpha`, `s_bak`, `y_bak`, `buffer_y` are kept for layout
 *      compatibility; the compact-form step does not need them.
 *
 *  Workspaces MUST come from initialize_*(): the library keeps private state
 *  (Gram matrices, streams, staging buffers) next to the public struct.  A struct
 *  assembled by hand is answered with task = invalid_input and -1000.
 *
 *  Device / stream / multi-GPU / bundled-callback extensions: stochqn_b200.h.
 *
 *  Precision is selected at compile time exactly like the reference
 *  (include/stochqn.h:62-76): default double, -DUSE_FLOAT for float.  Two shared
 *  libraries are built: libstochqn_b200_f64.so and libstochqn_b200_f32.so.
 */
#ifndef STOCHQN_INCLUDE
#define STOCHQN_INCLUDE

#include <stddef.h>

#if defined(USE_DOUBLE) || !defined(USE_FLOAT)
    #define real_t double
#else
    #define real_t float
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ---- correction-pair ring buffer (reference include/stochqn.h:86-99) ---------------
   s_mem / y_mem: row-major [mem_size][n] in HBM.  mem_st_ix = slot the NEXT pair is
   written to; mem_used = number of valid pairs.  y_reg, min_curvature and upd_freq are
   read on every call and may be changed between calls, as in the reference. */
typedef struct {
    real_t *s_mem;
    real_t *y_mem;
    real_t *buffer_rho;
    real_t *buffer_alpha;
    real_t *s_bak;
    real_t *y_bak;
    size_t mem_size;
    size_t mem_used;
    size_t mem_st_ix;
    size_t upd_freq;
    real_t y_reg;
    real_t min_curvature;
} bfgs_mem;

/* ---- empirical-Fisher ring of raw gradients (reference include/stochqn.h:101-107) -- */
typedef struct {
    real_t *F;
    real_t *buffer_y;
    size_t mem_size;
    size_t mem_used;
    size_t mem_st_ix;
} fisher_mem;

/* ---- per-optimizer workspaces (reference include/stochqn.h:109-151) ---------------- */
typedef struct {
    bfgs_mem *bfgs_memory;
    real_t *grad_prev;
    real_t hess_init;
    size_t niter;
    int section;
    int nthreads;
    int check_nan;
    int n;
} workspace_oLBFGS;

typedef struct {
    bfgs_mem *bfgs_memory;
    real_t *grad_prev;
    real_t *x_sum;
    real_t *x_avg_prev;
    int use_grad_diff;
    size_t niter;
    int section;
    int nthreads;
    int check_nan;
    int n;
} workspace_SQN;

typedef struct {
    bfgs_mem *bfgs_memory;
    fisher_mem *fisher_memory;
    real_t *H0;
    real_t *grad_prev;
    real_t *x_sum;
    real_t *x_avg_prev;
    real_t *grad_sum_sq;
    real_t f_prev;
    real_t max_incr;
    real_t scal_reg;
    real_t rmsprop_weight;
    int use_grad_diff;
    size_t niter;
    int section;
    int nthreads;
    int check_nan;
    int n;
} workspace_adaQN;

/* ---- request / status codes (reference include/stochqn.h:268-291), values bit-exact - */
typedef enum task_enum {
    calc_grad = 101,
    calc_grad_same_batch = 102,
    calc_grad_big_batch = 103,
    calc_hess_vec = 104,
    calc_fun_val_batch = 105,
    invalid_input = 100
} task_enum;

typedef enum info_enum {
    func_increased = 201,
    curvature_too_small = 202,
    search_direction_was_nan = 203,
    no_problems_encountered = 200
} info_enum;

typedef enum iter_status {
    did_not_update_x = 0,
    updated_x = 1,
    received_invalid_input = -1000
} iter_status;

/* ---- construction / destruction (reference include/stochqn.h:227-238) --------------
   Allocate the workspace and all of its HBM buffers on the current CUDA device
   (see stochqn_b200_set_device).  Return NULL (after a message on stderr) when host or
   device memory cannot be had - the reference's only error channel for allocation.
   `nthreads` is accepted and stored for compatibility; it has no effect on the GPU.
   LIMIT of this build (the reference has none): 1 <= mem_size <= 32 (STOCHQN_B200_MAX_MEM_SIZE) - the m x m
   compact-form solve runs in one warp.  A larger mem_size is refused like an allocation failure: message on
   stderr, NULL returned (stochqn_b200_last_error() carries the text). */
#define STOCHQN_B200_MAX_MEM_SIZE 32
workspace_oLBFGS* initialize_oLBFGS(const int n, const size_t mem_size, const real_t hess_init, const real_t y_reg,
    const real_t min_curvature, const int check_nan, const int nthreads);
void dealloc_oLBFGS(workspace_oLBFGS *oLBFGS);

workspace_SQN* initialize_SQN(const int n, const size_t mem_size, const size_t bfgs_upd_freq, const real_t min_curvature,
    const int use_grad_diff, const real_t y_reg, const int check_nan, const int nthreads);
void dealloc_SQN(workspace_SQN *SQN);

workspace_adaQN* initialize_adaQN(const int n, const size_t mem_size, const size_t fisher_size, const size_t bfgs_upd_freq,
    const real_t max_incr, const real_t min_curvature, const real_t scal_reg, const real_t rmsprop_weight,
    const int use_grad_diff, const real_t y_reg, const int check_nan, const int nthreads);
void dealloc_adaQN(workspace_adaQN *adaQN);

/* Externally visible in the reference although not prototyped there
   (src/stochqn.c:300-360); kept so that anything linking to them still resolves. */
bfgs_mem* initialize_bfgs_mem(const size_t mem_size, const int n, const real_t min_curvature, const real_t y_reg, const size_t upd_freq);
void dealloc_bfgs_mem(bfgs_mem *bfgs_memory);
fisher_mem* initialize_fisher_mem(const size_t mem_size, const int n);
void dealloc_fisher_mem(fisher_mem *fisher_memory);

/* ---- free-mode request loop (reference include/stochqn.h:381-383) -------------------
   Each call advances the optimizer's state machine as far as it can, then asks the
   caller for one calculation: `*task` says which (gradient on a new batch / on the same
   batch / on a big batch, Hessian-vector product, objective value), `*req` where to
   evaluate it, `*req_vec` (SQN) which vector to multiply.  The caller puts the answer
   in `grad` / `hess_vec` / `f` and calls again.  Returns 1 when `x` was updated, 0 when
   it was not, -1000 on an invalid workspace.  `grad` is overwritten.  On return all
   device work of the call has completed: `*task`, `*iter_info`, the counters in the
   workspace and the contents of `x` are final. */
int run_oLBFGS(real_t step_size, real_t x[], real_t grad[], real_t **req, task_enum *task, workspace_oLBFGS *oLBFGS, info_enum *iter_info);
int run_SQN(real_t step_size, real_t x[], real_t grad[], real_t hess_vec[], real_t **req, real_t **req_vec, task_enum *task, workspace_SQN *SQN, info_enum *iter_info);
int run_adaQN(real_t step_size, real_t x[], real_t f, real_t grad[], real_t **req, task_enum *task, workspace_adaQN *adaQN, info_enum *iter_info);

#ifdef __cplusplus
}
#endif

/* ---- C++ RAII front-ends, same class and method names as the reference's
        (include/stochqn.h:400-508) ------------------------------------------------- */
#ifdef __cplusplus
#include <new>

class oLBFGS
{
public:
    workspace_oLBFGS *workspace;
    task_enum task;
    info_enum info;
    iter_status status;
    real_t *req;

    oLBFGS(const int n, const size_t mem_size = 10, const real_t hess_init = 0, const real_t y_reg = 0,
           const real_t min_curvature = 0, const int check_nan = 1, const int nthreads = 1)
        : workspace(initialize_oLBFGS(n, mem_size, hess_init, y_reg, min_curvature, check_nan, nthreads)),
          task(calc_grad), info(no_problems_encountered), status(did_not_update_x), req(NULL)
    {
        if (!workspace) throw std::bad_alloc();
    }
    ~oLBFGS() { if (workspace) dealloc_oLBFGS(workspace); }

    iter_status run(real_t step_size, real_t x[], real_t grad[])
    {
        return (iter_status) run_oLBFGS(step_size, x, grad, &req, &task, workspace, &info);
    }
    task_enum get_task()      { return task; }
    info_enum get_iter_info() { return info; }
    size_t    get_n_iter()    { return workspace->niter; }
    real_t*   get_req()       { return req; }
};

class SQN
{
public:
    workspace_SQN *workspace;
    task_enum task;
    info_enum info;
    iter_status status;
    real_t *req;
    real_t *req_vec;

    SQN(const int n, const size_t mem_size = 10, const size_t bfgs_upd_freq = 10,
        const real_t min_curvature = 1e-4, const int use_grad_diff = 0, const real_t y_reg = 0,
        const int check_nan = 1, const int nthreads = 1)
        : workspace(initialize_SQN(n, mem_size, bfgs_upd_freq, min_curvature, use_grad_diff, y_reg, check_nan, nthreads)),
          task(calc_grad), info(no_problems_encountered), status(did_not_update_x), req(NULL), req_vec(NULL)
    {
        if (!workspace) throw std::bad_alloc();
    }
    ~SQN() { if (workspace) dealloc_SQN(workspace); }

    iter_status run(real_t step_size, real_t x[], real_t grad[], real_t hess_vec[])
    {
        return (iter_status) run_SQN(step_size, x, grad, hess_vec, &req, &req_vec, &task, workspace, &info);
    }
    task_enum get_task()      { return task; }
    info_enum get_iter_info() { return info; }
    size_t    get_n_iter()    { return workspace->niter; }
    real_t*   get_req()       { return req; }
    real_t*   get_req_vec()   { return req_vec; }
};

class adaQN
{
public:
    workspace_adaQN *workspace;
    task_enum task;
    info_enum info;
    iter_status status;
    real_t *req;

    adaQN(const int n, const size_t mem_size = 10, const size_t fisher_size = 100,
          const size_t bfgs_upd_freq = 10, const real_t max_incr = 1.01, const real_t min_curvature = 1e-4,
          const real_t scal_reg = 1e-4, const real_t rmsprop_weight = 0.9, const int use_grad_diff = 0,
          const real_t y_reg = 0, const int check_nan = 1, const int nthreads = 1)
        : workspace(initialize_adaQN(n, mem_size, fisher_size, bfgs_upd_freq, max_incr, min_curvature,
                                     scal_reg, rmsprop_weight, use_grad_diff, y_reg, check_nan, nthreads)),
          task(calc_grad), info(no_problems_encountered), status(did_not_update_x), req(NULL)
    {
        if (!workspace) throw std::bad_alloc();
    }
    ~adaQN() { if (workspace) dealloc_adaQN(workspace); }

    iter_status run(real_t step_size, real_t x[], real_t f, real_t grad[])
    {
        return (iter_status) run_adaQN(step_size, x, f, grad, &req, &task, workspace, &info);
    }
    task_enum get_task()      { return task; }
    info_enum get_iter_info() { return info; }
    size_t    get_n_iter()    { return workspace->niter; }
    real_t*   get_req()       { return req; }
};
#endif /* __cplusplus */

#endif /* STOCHQN_INCLUDE */
