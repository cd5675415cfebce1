/*  Rwrapper_b200.c - R .Call glue for the B200 build of the stochQN C ABI.
 *
 *  Replaces the reference's src/Rwrapper.c:98-229 (r_run_oLBFGS / r_run_SQN / r_run_adaQN and their struct
 *  re-assembly helpers, Rwrapper.c:17-96) and R/allocators.R.  The reference passes 19-33 R vectors per call and
 *  rebuilds the C structs on the stack from them (Rwrapper.c:106-110); here the arrays live in GPU memory inside a
 *  workspace created by initialize_*(), so R holds ONE external pointer per optimizer and the counters the R layer
 *  reads back after each call (niter, section, mem_used, mem_st_ix, Fisher counters, f_prev - Rwrapper.c:117-123,
 *  149-156, 185-194) are copied out of the public struct exactly as before.
 *
 *  x / grad / hess_vec are ordinary R numeric vectors (host memory): the library stages them through the device
 *  (compatibility mode) and `*req` / `*req_vec` point at host mirrors, copied into the pre-allocated R vectors as the
 *  reference does (Rwrapper.c:123).
 *
 *  Not compiled in this repository's image (no R); tests/test_r_binding_syntax.py compiles it against stub headers
 *  that declare the handful of R API entry points used here.
 *
 *  Link: PKG_LIBS = -L<...>/stochqn_b200/lib -lstochqn_b200_f64 (R numerics are double).
 */
#include <string.h>
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include "stochqn.h"
#include "stochqn_b200.h"

enum { KIND_OLBFGS = 1, KIND_SQN = 2, KIND_ADAQN = 3 };

typedef struct {
    int kind;
    void *ws;
} b200_handle;

static void release(b200_handle *h)
{
    if (!h || !h->ws) return;
    if (h->kind == KIND_OLBFGS) dealloc_oLBFGS((workspace_oLBFGS*) h->ws);
    else if (h->kind == KIND_SQN) dealloc_SQN((workspace_SQN*) h->ws);
    else dealloc_adaQN((workspace_adaQN*) h->ws);
    h->ws = NULL;
}

static void handle_finalizer(SEXP ptr)
{
    b200_handle *h = (b200_handle*) R_ExternalPtrAddr(ptr);
    if (!h) return;
    release(h);
    R_Free(h);
    R_ClearExternalPtr(ptr);
}

static SEXP make_handle(int kind, void *ws, const char *what)
{
    if (!ws) error("Could not create the %s workspace on the GPU: %s", what, stochqn_b200_last_error());
    b200_handle *h = R_Calloc(1, b200_handle);
    h->kind = kind;
    h->ws = ws;
    SEXP ptr = PROTECT(R_MakeExternalPtr(h, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(ptr, handle_finalizer, TRUE);
    UNPROTECT(1);
    return ptr;
}

static void* get_ws(SEXP ptr, int kind)
{
    b200_handle *h = (b200_handle*) R_ExternalPtrAddr(ptr);
    if (!h || !h->ws || h->kind != kind) error("Invalid or released optimizer workspace.");
    return h->ws;
}

/* ---- construction (reference: R/allocators.R create.r.oLBFGS / create.r.SQN / create.r.adaQN) ------------------- */
SEXP r_b200_init_oLBFGS(SEXP n, SEXP mem_size, SEXP hess_init, SEXP y_reg, SEXP min_curvature, SEXP check_nan, SEXP nthreads)
{
    return make_handle(KIND_OLBFGS,
                       initialize_oLBFGS(asInteger(n), (size_t) asInteger(mem_size), asReal(hess_init), asReal(y_reg),
                                         asReal(min_curvature), asInteger(check_nan), asInteger(nthreads)), "oLBFGS");
}

SEXP r_b200_init_SQN(SEXP n, SEXP mem_size, SEXP bfgs_upd_freq, SEXP min_curvature, SEXP use_grad_diff, SEXP y_reg,
                     SEXP check_nan, SEXP nthreads)
{
    return make_handle(KIND_SQN,
                       initialize_SQN(asInteger(n), (size_t) asInteger(mem_size), (size_t) asInteger(bfgs_upd_freq),
                                      asReal(min_curvature), asInteger(use_grad_diff), asReal(y_reg), asInteger(check_nan),
                                      asInteger(nthreads)), "SQN");
}

SEXP r_b200_init_adaQN(SEXP n, SEXP mem_size, SEXP fisher_size, SEXP bfgs_upd_freq, SEXP max_incr, SEXP min_curvature,
                       SEXP scal_reg, SEXP rmsprop_weight, SEXP use_grad_diff, SEXP y_reg, SEXP check_nan, SEXP nthreads)
{
    return make_handle(KIND_ADAQN,
                       initialize_adaQN(asInteger(n), (size_t) asInteger(mem_size), (size_t) asInteger(fisher_size),
                                        (size_t) asInteger(bfgs_upd_freq), asReal(max_incr), asReal(min_curvature),
                                        asReal(scal_reg), asReal(rmsprop_weight), asInteger(use_grad_diff), asReal(y_reg),
                                        asInteger(check_nan), asInteger(nthreads)), "adaQN");
}

SEXP r_b200_release(SEXP ptr)
{
    b200_handle *h = (b200_handle*) R_ExternalPtrAddr(ptr);
    release(h);
    return R_NilValue;
}

/* ---- the request loop: same out-parameters as the reference's r_run_* (pre-allocated R vectors, written in place) -- */
SEXP r_b200_run_oLBFGS(SEXP ws_ptr, SEXP x, SEXP grad, SEXP step_size,
                       SEXP niter, SEXP section, SEXP mem_used, SEXP mem_st_ix,
                       SEXP x_changed, SEXP req_R, SEXP task_R, SEXP iter_info_R)
{
    workspace_oLBFGS *ws = (workspace_oLBFGS*) get_ws(ws_ptr, KIND_OLBFGS);
    info_enum iter_info;
    task_enum task;
    double *req = NULL;
    INTEGER(x_changed)[0] = run_oLBFGS(REAL(step_size)[0], REAL(x), REAL(grad), &req, &task, ws, &iter_info);
    INTEGER(niter)[0] = (int) ws->niter;
    INTEGER(section)[0] = ws->section;
    INTEGER(mem_used)[0] = (int) ws->bfgs_memory->mem_used;
    INTEGER(mem_st_ix)[0] = (int) ws->bfgs_memory->mem_st_ix;
    INTEGER(task_R)[0] = (int) task;
    INTEGER(iter_info_R)[0] = (int) iter_info;
    if (req) memcpy(REAL(req_R), req, (size_t) ws->n * sizeof(double));
    return R_NilValue;
}

SEXP r_b200_run_SQN(SEXP ws_ptr, SEXP x, SEXP grad, SEXP hess_vec, SEXP step_size,
                    SEXP niter, SEXP section, SEXP mem_used, SEXP mem_st_ix,
                    SEXP x_changed, SEXP req_R, SEXP req_vec_R, SEXP task_R, SEXP iter_info_R)
{
    workspace_SQN *ws = (workspace_SQN*) get_ws(ws_ptr, KIND_SQN);
    info_enum iter_info;
    task_enum task;
    double *req = NULL, *req_vec = NULL;
    INTEGER(x_changed)[0] = run_SQN(REAL(step_size)[0], REAL(x), REAL(grad), REAL(hess_vec), &req, &req_vec, &task, ws, &iter_info);
    INTEGER(niter)[0] = (int) ws->niter;
    INTEGER(section)[0] = ws->section;
    INTEGER(mem_used)[0] = (int) ws->bfgs_memory->mem_used;
    INTEGER(mem_st_ix)[0] = (int) ws->bfgs_memory->mem_st_ix;
    INTEGER(task_R)[0] = (int) task;
    INTEGER(iter_info_R)[0] = (int) iter_info;
    if (req) memcpy(REAL(req_R), req, (size_t) ws->n * sizeof(double));
    if (task == calc_hess_vec && req_vec) memcpy(REAL(req_vec_R), req_vec, (size_t) ws->n * sizeof(double));
    return R_NilValue;
}

SEXP r_b200_run_adaQN(SEXP ws_ptr, SEXP x, SEXP f, SEXP grad, SEXP step_size,
                      SEXP niter, SEXP section, SEXP mem_used, SEXP mem_st_ix,
                      SEXP fisher_used, SEXP fisher_st_ix, SEXP f_prev,
                      SEXP x_changed, SEXP req_R, SEXP task_R, SEXP iter_info_R)
{
    workspace_adaQN *ws = (workspace_adaQN*) get_ws(ws_ptr, KIND_ADAQN);
    info_enum iter_info;
    task_enum task;
    double *req = NULL;
    INTEGER(x_changed)[0] = run_adaQN(REAL(step_size)[0], REAL(x), REAL(f)[0], REAL(grad), &req, &task, ws, &iter_info);
    INTEGER(niter)[0] = (int) ws->niter;
    INTEGER(section)[0] = ws->section;
    INTEGER(mem_used)[0] = (int) ws->bfgs_memory->mem_used;
    INTEGER(mem_st_ix)[0] = (int) ws->bfgs_memory->mem_st_ix;
    if (ws->fisher_memory) {
        INTEGER(fisher_used)[0] = (int) ws->fisher_memory->mem_used;
        INTEGER(fisher_st_ix)[0] = (int) ws->fisher_memory->mem_st_ix;
    }
    REAL(f_prev)[0] = ws->f_prev;
    INTEGER(task_R)[0] = (int) task;
    INTEGER(iter_info_R)[0] = (int) iter_info;
    if (req) memcpy(REAL(req_R), req, (size_t) ws->n * sizeof(double));
    return R_NilValue;
}

/* tunables the reference documents as modifiable between calls (include/stochqn.h:163-167) */
SEXP r_b200_set_f_prev(SEXP ws_ptr, SEXP value)
{
    workspace_adaQN *ws = (workspace_adaQN*) get_ws(ws_ptr, KIND_ADAQN);
    ws->f_prev = asReal(value);
    return R_NilValue;
}

static const R_CallMethodDef call_methods[] = {
    {"r_b200_init_oLBFGS", (DL_FUNC) &r_b200_init_oLBFGS, 7},
    {"r_b200_init_SQN", (DL_FUNC) &r_b200_init_SQN, 8},
    {"r_b200_init_adaQN", (DL_FUNC) &r_b200_init_adaQN, 12},
    {"r_b200_release", (DL_FUNC) &r_b200_release, 1},
    {"r_b200_run_oLBFGS", (DL_FUNC) &r_b200_run_oLBFGS, 12},
    {"r_b200_run_SQN", (DL_FUNC) &r_b200_run_SQN, 14},
    {"r_b200_run_adaQN", (DL_FUNC) &r_b200_run_adaQN, 16},
    {"r_b200_set_f_prev", (DL_FUNC) &r_b200_set_f_prev, 2},
    {NULL, NULL, 0}
};

void R_init_stochQNb200(DllInfo *info)
{
    R_registerRoutines(info, NULL, call_methods, NULL, NULL);
    R_useDynamicSymbols(info, TRUE);
}
