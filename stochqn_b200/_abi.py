"""ctypes mirror of ``include/stochqn.h`` - struct layouts, enum values, prototypes.

The same binding serves any library that exports the stochQN C ABI (reference
include/stochqn.h:86-151, 227-238, 268-291, 381-383): the CUDA libraries built from
``stochqn_b200/csrc`` (product) and, in the test-suite only, the CPU oracle compiled
from the reference sources.  Nothing here computes anything.
"""
from __future__ import annotations

import ctypes as C

# task_enum (reference include/stochqn.h:268-275)
CALC_GRAD = 101
CALC_GRAD_SAME_BATCH = 102
CALC_GRAD_BIG_BATCH = 103
CALC_HESS_VEC = 104
CALC_FUN_VAL_BATCH = 105
INVALID_INPUT = 100
TASK_NAMES = {
    CALC_GRAD: "calc_grad",
    CALC_GRAD_SAME_BATCH: "calc_grad_same_batch",
    CALC_GRAD_BIG_BATCH: "calc_grad_big_batch",
    CALC_HESS_VEC: "calc_hess_vec",
    CALC_FUN_VAL_BATCH: "calc_fun_val_batch",
    INVALID_INPUT: "invalid_input",
}
# info_enum (reference include/stochqn.h:279-284)
NO_PROBLEMS = 200
FUNC_INCREASED = 201
CURVATURE_TOO_SMALL = 202
SEARCH_DIRECTION_WAS_NAN = 203
INFO_NAMES = {
    NO_PROBLEMS: "no_problems_encountered",
    FUNC_INCREASED: "func_increased",
    CURVATURE_TOO_SMALL: "curvature_too_small",
    SEARCH_DIRECTION_WAS_NAN: "search_direction_was_nan",
}
# iter_status (reference include/stochqn.h:291)
DID_NOT_UPDATE_X = 0
UPDATED_X = 1
RECEIVED_INVALID_INPUT = -1000


def make_structs(real):
    """Struct mirrors for one precision (``ctypes.c_double`` or ``ctypes.c_float``)."""
    P = C.POINTER(real)

    class bfgs_mem(C.Structure):
        _fields_ = [
            ("s_mem", P), ("y_mem", P), ("buffer_rho", P), ("buffer_alpha", P),
            ("s_bak", P), ("y_bak", P),
            ("mem_size", C.c_size_t), ("mem_used", C.c_size_t), ("mem_st_ix", C.c_size_t),
            ("upd_freq", C.c_size_t), ("y_reg", real), ("min_curvature", real),
        ]

    class fisher_mem(C.Structure):
        _fields_ = [
            ("F", P), ("buffer_y", P),
            ("mem_size", C.c_size_t), ("mem_used", C.c_size_t), ("mem_st_ix", C.c_size_t),
        ]

    class workspace_oLBFGS(C.Structure):
        _fields_ = [
            ("bfgs_memory", C.POINTER(bfgs_mem)), ("grad_prev", P), ("hess_init", real),
            ("niter", C.c_size_t), ("section", C.c_int), ("nthreads", C.c_int),
            ("check_nan", C.c_int), ("n", C.c_int),
        ]

    class workspace_SQN(C.Structure):
        _fields_ = [
            ("bfgs_memory", C.POINTER(bfgs_mem)), ("grad_prev", P), ("x_sum", P), ("x_avg_prev", P),
            ("use_grad_diff", C.c_int), ("niter", C.c_size_t), ("section", C.c_int),
            ("nthreads", C.c_int), ("check_nan", C.c_int), ("n", C.c_int),
        ]

    class workspace_adaQN(C.Structure):
        _fields_ = [
            ("bfgs_memory", C.POINTER(bfgs_mem)), ("fisher_memory", C.POINTER(fisher_mem)),
            ("H0", P), ("grad_prev", P), ("x_sum", P), ("x_avg_prev", P), ("grad_sum_sq", P),
            ("f_prev", real), ("max_incr", real), ("scal_reg", real), ("rmsprop_weight", real),
            ("use_grad_diff", C.c_int), ("niter", C.c_size_t), ("section", C.c_int),
            ("nthreads", C.c_int), ("check_nan", C.c_int), ("n", C.c_int),
        ]

    return dict(bfgs_mem=bfgs_mem, fisher_mem=fisher_mem, workspace_oLBFGS=workspace_oLBFGS,
                workspace_SQN=workspace_SQN, workspace_adaQN=workspace_adaQN)


class StochqnABI:
    """The nine reference entry points of one loaded library, with argtypes set.

    Array arguments are declared ``c_void_p`` so that raw device addresses
    (``tensor.data_ptr()``) and host arrays (``ndarray.ctypes.data``) pass alike.
    """

    def __init__(self, lib: C.CDLL, real):
        self.lib = lib
        self.real = real
        self.structs = make_structs(real)
        S = self.structs
        vp, sz, ci = C.c_void_p, C.c_size_t, C.c_int
        PP = C.POINTER(C.c_void_p)
        PI = C.POINTER(C.c_int)

        lib.initialize_oLBFGS.argtypes = [ci, sz, real, real, real, ci, ci]
        lib.initialize_oLBFGS.restype = C.POINTER(S["workspace_oLBFGS"])
        lib.dealloc_oLBFGS.argtypes = [C.POINTER(S["workspace_oLBFGS"])]
        lib.dealloc_oLBFGS.restype = None

        lib.initialize_SQN.argtypes = [ci, sz, sz, real, ci, real, ci, ci]
        lib.initialize_SQN.restype = C.POINTER(S["workspace_SQN"])
        lib.dealloc_SQN.argtypes = [C.POINTER(S["workspace_SQN"])]
        lib.dealloc_SQN.restype = None

        lib.initialize_adaQN.argtypes = [ci, sz, sz, sz, real, real, real, real, ci, real, ci, ci]
        lib.initialize_adaQN.restype = C.POINTER(S["workspace_adaQN"])
        lib.dealloc_adaQN.argtypes = [C.POINTER(S["workspace_adaQN"])]
        lib.dealloc_adaQN.restype = None

        lib.run_oLBFGS.argtypes = [real, vp, vp, PP, PI, C.POINTER(S["workspace_oLBFGS"]), PI]
        lib.run_oLBFGS.restype = ci
        lib.run_SQN.argtypes = [real, vp, vp, vp, PP, PP, PI, C.POINTER(S["workspace_SQN"]), PI]
        lib.run_SQN.restype = ci
        lib.run_adaQN.argtypes = [real, vp, real, vp, PP, PI, C.POINTER(S["workspace_adaQN"]), PI]
        lib.run_adaQN.restype = ci

    REFERENCE_SYMBOLS = (
        "initialize_oLBFGS", "dealloc_oLBFGS", "initialize_SQN", "dealloc_SQN",
        "initialize_adaQN", "dealloc_adaQN", "run_oLBFGS", "run_SQN", "run_adaQN",
        "initialize_bfgs_mem", "dealloc_bfgs_mem", "initialize_fisher_mem", "dealloc_fisher_mem",
    )
