# cython: language_level=3, boundscheck=False, wraparound=False
"""Cython binding of the stochQN C ABI on the B200 library (SURVEY section 8(f), row 1).

Replaces the reference's stochqn/pywrapper.pxi:1-207 (+ wrapper_double.pyx / wrapper_float.pyx): the struct
re-assembly from holder objects (pywrapper.pxi:89-159, once per call) disappears - the holders keep one workspace
address made by ``py_init_*`` - and the ``py_run_*`` functions return the same tuples as pywrapper.pxi:161-207.
``x`` / ``grad`` / ``hess_vec`` are NumPy arrays (host memory, staged by the library) exactly as in the reference;
``py_run_*_ptr`` take raw device addresses (``tensor.data_ptr()``) for GPU-resident callers.

The repository's own Python layer (stochqn_b200/optimizers.py) binds the same ABI through ctypes and is what the test
suite exercises; this file is for maintainers who keep the reference's Cython layer.  Built for the double library
(compile with -DUSE_FLOAT and link libstochqn_b200_f32 for the float twin, as the reference does with its two .pyx).
"""
import numpy as np
cimport numpy as np
from libc.stdint cimport uintptr_t

ctypedef double real_t

cdef extern from "stochqn.h":
    ctypedef struct bfgs_mem:
        size_t mem_size
        size_t mem_used
        size_t mem_st_ix
        size_t upd_freq
        real_t y_reg
        real_t min_curvature
    ctypedef struct fisher_mem:
        size_t mem_size
        size_t mem_used
        size_t mem_st_ix
    ctypedef struct workspace_oLBFGS:
        bfgs_mem *bfgs_memory
        size_t niter
        int section
        int n
    ctypedef struct workspace_SQN:
        bfgs_mem *bfgs_memory
        size_t niter
        int section
        int n
    ctypedef struct workspace_adaQN:
        bfgs_mem *bfgs_memory
        fisher_mem *fisher_memory
        real_t f_prev
        size_t niter
        int section
        int n
    ctypedef enum task_enum:
        calc_grad = 101
        calc_grad_same_batch = 102
        calc_grad_big_batch = 103
        calc_hess_vec = 104
        calc_fun_val_batch = 105
        invalid_input = 100
    ctypedef enum info_enum:
        func_increased = 201
        curvature_too_small = 202
        search_direction_was_nan = 203
        no_problems_encountered = 200
    workspace_oLBFGS* initialize_oLBFGS(int n, size_t mem_size, real_t hess_init, real_t y_reg, real_t min_curvature,
                                        int check_nan, int nthreads)
    workspace_SQN* initialize_SQN(int n, size_t mem_size, size_t bfgs_upd_freq, real_t min_curvature, int use_grad_diff,
                                  real_t y_reg, int check_nan, int nthreads)
    workspace_adaQN* initialize_adaQN(int n, size_t mem_size, size_t fisher_size, size_t bfgs_upd_freq, real_t max_incr,
                                      real_t min_curvature, real_t scal_reg, real_t rmsprop_weight, int use_grad_diff,
                                      real_t y_reg, int check_nan, int nthreads)
    void dealloc_oLBFGS(workspace_oLBFGS *ws)
    void dealloc_SQN(workspace_SQN *ws)
    void dealloc_adaQN(workspace_adaQN *ws)
    int run_oLBFGS(real_t step_size, real_t *x, real_t *grad, real_t **req, task_enum *task, workspace_oLBFGS *ws,
                   info_enum *iter_info)
    int run_SQN(real_t step_size, real_t *x, real_t *grad, real_t *hess_vec, real_t **req, real_t **req_vec,
                task_enum *task, workspace_SQN *ws, info_enum *iter_info)
    int run_adaQN(real_t step_size, real_t *x, real_t f, real_t *grad, real_t **req, task_enum *task,
                  workspace_adaQN *ws, info_enum *iter_info)

cdef extern from "stochqn_b200.h":
    const char* stochqn_b200_last_error()


cdef object _host_view(real_t *p, Py_ssize_t n):
    """NumPy view of a host mirror handed out through *req (no copy), as pywrapper.pxi:172 does."""
    if p == NULL:
        return None
    cdef real_t[::1] mv = <real_t[:n]> p
    return np.asarray(mv)


# ---- construction / destruction: the holders of stochqn/_optimizers.py:791-879 keep only these addresses ------------
def py_init_oLBFGS(int n, size_t mem_size, real_t hess_init, real_t y_reg, real_t min_curvature, int check_nan, int nthreads):
    cdef workspace_oLBFGS *ws = initialize_oLBFGS(n, mem_size, hess_init, y_reg, min_curvature, check_nan, nthreads)
    if ws == NULL:
        raise MemoryError("initialize_oLBFGS failed: " + stochqn_b200_last_error().decode())
    return <uintptr_t> ws

def py_init_SQN(int n, size_t mem_size, size_t bfgs_upd_freq, real_t min_curvature, int use_grad_diff, real_t y_reg,
                int check_nan, int nthreads):
    cdef workspace_SQN *ws = initialize_SQN(n, mem_size, bfgs_upd_freq, min_curvature, use_grad_diff, y_reg, check_nan, nthreads)
    if ws == NULL:
        raise MemoryError("initialize_SQN failed: " + stochqn_b200_last_error().decode())
    return <uintptr_t> ws

def py_init_adaQN(int n, size_t mem_size, size_t fisher_size, size_t bfgs_upd_freq, real_t max_incr, real_t min_curvature,
                  real_t scal_reg, real_t rmsprop_weight, int use_grad_diff, real_t y_reg, int check_nan, int nthreads):
    cdef workspace_adaQN *ws = initialize_adaQN(n, mem_size, fisher_size, bfgs_upd_freq, max_incr, min_curvature, scal_reg,
                                                rmsprop_weight, use_grad_diff, y_reg, check_nan, nthreads)
    if ws == NULL:
        raise MemoryError("initialize_adaQN failed: " + stochqn_b200_last_error().decode())
    return <uintptr_t> ws

def py_free_oLBFGS(uintptr_t ws_addr):
    dealloc_oLBFGS(<workspace_oLBFGS*> ws_addr)

def py_free_SQN(uintptr_t ws_addr):
    dealloc_SQN(<workspace_SQN*> ws_addr)

def py_free_adaQN(uintptr_t ws_addr):
    dealloc_adaQN(<workspace_adaQN*> ws_addr)


# ---- the request loop: same returned tuples as pywrapper.pxi:161-207 ---------------------------------------------------
def py_run_oLBFGS(uintptr_t ws_addr, np.ndarray[real_t, ndim=1, mode="c"] x, np.ndarray[real_t, ndim=1, mode="c"] grad,
                  real_t step_size):
    cdef workspace_oLBFGS *ws = <workspace_oLBFGS*> ws_addr
    cdef real_t *req = NULL
    cdef task_enum task
    cdef info_enum iter_info
    cdef int x_changed = run_oLBFGS(step_size, &x[0], &grad[0], &req, &task, ws, &iter_info)
    if x_changed == -1000:
        raise ValueError("run_oLBFGS received invalid input: " + stochqn_b200_last_error().decode())
    return (x_changed, ws.niter, ws.section, ws.bfgs_memory.mem_used, ws.bfgs_memory.mem_st_ix, <int> task, <int> iter_info,
            x if req == &x[0] else _host_view(req, x.shape[0]))

def py_run_SQN(uintptr_t ws_addr, np.ndarray[real_t, ndim=1, mode="c"] x, real_t step_size,
               np.ndarray[real_t, ndim=1, mode="c"] grad, np.ndarray[real_t, ndim=1, mode="c"] hess_vec):
    cdef workspace_SQN *ws = <workspace_SQN*> ws_addr
    cdef real_t *req = NULL
    cdef real_t *req_vec = NULL
    cdef task_enum task
    cdef info_enum iter_info
    cdef int x_changed = run_SQN(step_size, &x[0], &grad[0], &hess_vec[0], &req, &req_vec, &task, ws, &iter_info)
    if x_changed == -1000:
        raise ValueError("run_SQN received invalid input: " + stochqn_b200_last_error().decode())
    return (x_changed, ws.niter, ws.section, ws.bfgs_memory.mem_used, ws.bfgs_memory.mem_st_ix, <int> task, <int> iter_info,
            x if req == &x[0] else _host_view(req, x.shape[0]),
            _host_view(req_vec, x.shape[0]) if task == calc_hess_vec else None)

def py_run_adaQN(uintptr_t ws_addr, np.ndarray[real_t, ndim=1, mode="c"] x, np.ndarray[real_t, ndim=1, mode="c"] grad,
                 real_t step_size, real_t f):
    cdef workspace_adaQN *ws = <workspace_adaQN*> ws_addr
    cdef real_t *req = NULL
    cdef task_enum task
    cdef info_enum iter_info
    cdef int x_changed = run_adaQN(step_size, &x[0], f, &grad[0], &req, &task, ws, &iter_info)
    if x_changed == -1000:
        raise ValueError("run_adaQN received invalid input: " + stochqn_b200_last_error().decode())
    cdef size_t f_used = ws.fisher_memory.mem_used if ws.fisher_memory != NULL else 0
    cdef size_t f_st = ws.fisher_memory.mem_st_ix if ws.fisher_memory != NULL else 0
    return (x_changed, ws.niter, ws.section, ws.bfgs_memory.mem_used, ws.bfgs_memory.mem_st_ix, f_used, f_st, ws.f_prev,
            <int> task, <int> iter_info, x if req == &x[0] else _host_view(req, x.shape[0]))


# ---- GPU-resident callers: raw device addresses in, raw device addresses out (no NumPy involved) ----------------------
def py_run_oLBFGS_ptr(uintptr_t ws_addr, uintptr_t x_dev, uintptr_t grad_dev, real_t step_size):
    cdef workspace_oLBFGS *ws = <workspace_oLBFGS*> ws_addr
    cdef real_t *req = NULL
    cdef task_enum task
    cdef info_enum iter_info
    cdef int x_changed = run_oLBFGS(step_size, <real_t*> x_dev, <real_t*> grad_dev, &req, &task, ws, &iter_info)
    return (x_changed, ws.niter, ws.section, ws.bfgs_memory.mem_used, ws.bfgs_memory.mem_st_ix, <int> task, <int> iter_info,
            <uintptr_t> req)
