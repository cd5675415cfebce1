"""Loader for the in-tree CUDA libraries (``stochqn_b200/lib/libstochqn_b200_{f64,f32}.so``).

There is no CPU path: a missing library raises, and the library itself fails loudly on a box
without a CUDA device.  The libraries are built by ``stochqn_b200/build.py`` (nvcc, sm_100a).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from ._abi import StochqnABI

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}

EXT_SYMBOLS = (
    "stochqn_b200_version", "stochqn_b200_real_bytes", "stochqn_b200_last_error", "stochqn_b200_launch_count",
    "stochqn_b200_set_stream", "stochqn_b200_set_option", "stochqn_b200_get_option", "stochqn_b200_debug_fit_trace", "stochqn_b200_debug_mn_trace", "stochqn_b200_get_stat", "stochqn_b200_row_stride",
    "stochqn_b200_comm_unique_id", "stochqn_b200_comm_init", "stochqn_b200_comm_destroy", "stochqn_b200_comm_uses_p2p",
    "stochqn_b200_set_comm",
    "stochqn_b200_allreduce_f64", "stochqn_b200_allreduce_real", "stochqn_b200_reduce_scatter_real", "stochqn_b200_all_gather_real",
    "stochqn_b200_rosenbrock_x0", "stochqn_b200_rosenbrock_grad", "stochqn_b200_rosenbrock_fun", "stochqn_b200_rosenbrock_halo",
    "stochqn_b200_rosenbrock_grad_sharded",
    "stochqn_b200_logistic_work_size", "stochqn_b200_csr_to_dense", "stochqn_b200_logistic_grad", "stochqn_b200_logistic_hess_vec",
    "stochqn_b200_logistic_loss", "stochqn_b200_export", "stochqn_b200_import",
    "stochqn_b200_logistic_sk_grad", "stochqn_b200_logistic_sk_hess_vec", "stochqn_b200_logistic_sk_loss",
    "stochqn_b200_multinomial_work_size", "stochqn_b200_multinomial_loss_grad", "stochqn_b200_multinomial_hess_vec",
    "stochqn_b200_gemm_tn", "stochqn_b200_fit_batch", "stochqn_b200_fit_batches", "stochqn_b200_multinomial_grad_reduce_scatter",
    "stochqn_b200_all_gather_p2p", "stochqn_b200_p2p_send_buffer", "stochqn_b200_reduce_scatter_p2p",
    "stochqn_b200_comm_init_inprocess", "stochqn_b200_comm_error",
)

OPT_GRAD_WRITEBACK = 1
OPT_TRUST_X_MIRROR = 2
OPT_PROFILE = 3
OPT_SYNC_RETURN = 4
OPT_ONE_LAUNCH_MAX_N = 5
OPT_DEVICE_LOOP_MAX_N = 6
OPT_FUSED_FIT = 7
STAT_K1_MS, STAT_K1_COUNT, STAT_K3_MS, STAT_K3_COUNT, STAT_K4_MS, STAT_K4_COUNT, STAT_LAST_BOUND = 1, 2, 3, 4, 5, 6, 7
STAT_EXACT_NORM_STEPS = 8
STAT_KA2_MS, STAT_KA2_COUNT = 9, 10
STAT_ONE_LAUNCH_STEPS = 11
STAT_DEVICE_LOOP_STEPS = 12
STAT_FUSED_FIT_STEPS = 13


def lib_path(dtype) -> str:
    tag = "f64" if np.dtype(dtype) == np.float64 else "f32"
    return os.path.join(_HERE, "lib", "libstochqn_b200_%s.so" % tag)


class HostState(C.Structure):
    """stochqn_b200_host_state (include/stochqn_b200.h)."""
    _fields_ = [(k, C.c_void_p) for k in
                ("s_mem", "y_mem", "grad_prev", "x_sum", "x_avg_prev", "grad_sum_sq", "F")]


class Rows(C.Structure):
    """stochqn_b200_rows (include/stochqn_b200.h): a row range of device-resident arrays."""
    _fields_ = [("X", C.c_void_p), ("ldx", C.c_longlong), ("y", C.c_void_p), ("ldy", C.c_longlong),
                ("sw", C.c_void_p), ("nrows", C.c_longlong)]


def model_struct(real):
    """stochqn_b200_model for the precision whose C real type is `real`."""
    class Model(C.Structure):
        _fields_ = [("model", C.c_int), ("fit_intercept", C.c_int), ("ncols", C.c_longlong), ("nclasses", C.c_longlong),
                    ("reg_param", real), ("work", C.c_void_p)]
    return Model


class FitReport(C.Structure):
    """stochqn_b200_fit_report."""
    _fields_ = [("calls", C.c_longlong), ("n_info", C.c_longlong * 4), ("last_info", C.c_int), ("x_changed", C.c_int),
                ("long_batch_used", C.c_int)]


def load(dtype=np.float64) -> StochqnABI:
    """Load (once) the library for `dtype` and declare every prototype."""
    key = np.dtype(dtype).name
    if key in _LIBS:
        return _LIBS[key]
    path = lib_path(dtype)
    if not os.path.exists(path):
        raise RuntimeError(
            "stochqn_b200: %s is missing - build it with `python -m stochqn_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback." % path)
    lib = C.CDLL(path)
    real = C.c_double if np.dtype(dtype) == np.float64 else C.c_float
    abi = StochqnABI(lib, real)
    vp, ll, ci, sz = C.c_void_p, C.c_longlong, C.c_int, C.c_size_t
    lib.stochqn_b200_version.restype = ci
    lib.stochqn_b200_real_bytes.restype = ci
    lib.stochqn_b200_last_error.restype = C.c_char_p
    lib.stochqn_b200_launch_count.restype = C.c_ulonglong
    lib.stochqn_b200_set_stream.argtypes = [vp, vp]
    lib.stochqn_b200_set_option.argtypes = [vp, ci, ll]
    lib.stochqn_b200_get_option.argtypes = [vp, ci]
    lib.stochqn_b200_get_option.restype = ll
    lib.stochqn_b200_debug_fit_trace.argtypes = [vp, vp]
    lib.stochqn_b200_debug_mn_trace.argtypes = [vp]
    lib.stochqn_b200_get_stat.argtypes = [vp, ci, C.POINTER(C.c_double)]
    lib.stochqn_b200_row_stride.argtypes = [vp]
    lib.stochqn_b200_row_stride.restype = sz
    lib.stochqn_b200_comm_unique_id.argtypes = [vp]
    lib.stochqn_b200_comm_init.argtypes = [vp, ci, ci, C.POINTER(vp)]
    lib.stochqn_b200_comm_destroy.argtypes = [vp]
    lib.stochqn_b200_comm_uses_p2p.argtypes = [vp]
    lib.stochqn_b200_set_comm.argtypes = [vp, vp, ll]
    lib.stochqn_b200_allreduce_f64.argtypes = [vp, vp, sz, vp]
    lib.stochqn_b200_allreduce_real.argtypes = [vp, vp, sz, vp]
    lib.stochqn_b200_reduce_scatter_real.argtypes = [vp, vp, vp, sz, vp]
    lib.stochqn_b200_all_gather_real.argtypes = [vp, vp, vp, sz, vp]
    lib.stochqn_b200_rosenbrock_x0.argtypes = [vp, ll, ll, vp]
    lib.stochqn_b200_rosenbrock_grad.argtypes = [vp, vp, ll, ll, ll, vp, vp]
    lib.stochqn_b200_rosenbrock_fun.argtypes = [vp, ll, ll, ll, vp, vp, vp]
    lib.stochqn_b200_rosenbrock_halo.argtypes = [vp, ll, ci, ci, vp, vp, vp, vp]
    lib.stochqn_b200_rosenbrock_grad_sharded.argtypes = [vp, vp, ll, ll, ll, ci, ci, vp, vp, vp, vp]
    lib.stochqn_b200_logistic_work_size.argtypes = [ll, ll]
    lib.stochqn_b200_logistic_work_size.restype = sz
    lib.stochqn_b200_csr_to_dense.argtypes = [vp, vp, vp, ll, ll, ll, vp, ll, vp, vp]
    lib.stochqn_b200_logistic_grad.argtypes = [vp, ll, vp, vp, ll, ll, vp, real, vp, vp, vp]
    lib.stochqn_b200_logistic_hess_vec.argtypes = [vp, ll, vp, vp, ll, ll, vp, vp, real, vp, vp, vp]
    lib.stochqn_b200_logistic_loss.argtypes = [vp, ll, vp, vp, ll, ll, vp, real, vp, vp, vp]
    lib.stochqn_b200_logistic_sk_grad.argtypes = [vp, ll, vp, vp, ll, ll, ci, vp, real, vp, vp, vp]
    lib.stochqn_b200_logistic_sk_hess_vec.argtypes = [vp, ll, vp, vp, ll, ll, ci, vp, vp, real, vp, vp, vp]
    lib.stochqn_b200_logistic_sk_loss.argtypes = [vp, ll, vp, vp, ll, ll, ci, vp, real, vp, vp, vp]
    lib.stochqn_b200_multinomial_work_size.argtypes = [ll, ll, ll]
    lib.stochqn_b200_multinomial_work_size.restype = sz
    lib.stochqn_b200_multinomial_loss_grad.argtypes = [vp, ll, vp, ll, vp, vp, ll, ll, ll, ci, vp, real, vp, vp, vp, vp]
    lib.stochqn_b200_multinomial_hess_vec.argtypes = [vp, ll, vp, ll, vp, vp, ll, ll, ll, ci, vp, vp, real, vp, vp, vp]
    lib.stochqn_b200_multinomial_grad_reduce_scatter.argtypes = [vp, vp, ll, vp, ll, vp, vp, ll, ll, ll, ci, vp, real, vp, ll, vp, vp]
    lib.stochqn_b200_all_gather_p2p.argtypes = [vp, vp, sz, C.POINTER(vp), vp]
    lib.stochqn_b200_p2p_send_buffer.argtypes = [vp, sz, C.POINTER(vp)]
    lib.stochqn_b200_reduce_scatter_p2p.argtypes = [vp, vp, vp, sz, vp]
    lib.stochqn_b200_comm_init_inprocess.argtypes = [ci, C.POINTER(vp)]
    lib.stochqn_b200_comm_error.argtypes = [vp]
    lib.stochqn_b200_gemm_tn.argtypes = [vp, ll, vp, ll, vp, ll, ci, ci, ci, vp]
    abi.Model = model_struct(real)
    lib.stochqn_b200_fit_batch.argtypes = [vp, vp, real, C.POINTER(abi.Model), C.POINTER(Rows), C.POINTER(Rows), C.POINTER(Rows),
                                           C.POINTER(ci), C.POINTER(vp), C.POINTER(vp), C.POINTER(FitReport)]
    lib.stochqn_b200_fit_batches.argtypes = [vp, vp, real, C.POINTER(abi.Model), C.POINTER(Rows), ll, ll, ll, C.POINTER(ll), C.POINTER(ll),
                                             C.POINTER(Rows), C.POINTER(ci), C.POINTER(vp), C.POINTER(vp), C.POINTER(FitReport)]
    lib.stochqn_b200_export.argtypes = [vp, C.POINTER(HostState)]
    lib.stochqn_b200_import.argtypes = [vp, C.POINTER(HostState)]
    for name in EXT_SYMBOLS:
        getattr(lib, name)        # AttributeError here = the library is stale: rebuild it
    if lib.stochqn_b200_real_bytes() != C.sizeof(real):
        raise RuntimeError("stochqn_b200: %s was built for another precision" % path)
    _LIBS[key] = abi
    return abi


def last_error(abi: StochqnABI) -> str:
    msg = abi.lib.stochqn_b200_last_error()
    return msg.decode() if msg else ""


def get_stat(abi: StochqnABI, ws, what: int) -> float:
    out = C.c_double()
    if abi.lib.stochqn_b200_get_stat(ws, what, C.byref(out)) != 0:
        raise RuntimeError(last_error(abi))
    return out.value


def launch_count() -> int:
    """Kernels launched so far by every loaded precision of the library."""
    return int(sum(a.lib.stochqn_b200_launch_count() for a in _LIBS.values()))
