"""TEST INFRASTRUCTURE ONLY - ctypes front-end to the reference C library in oracle/_ref.

``oracle/_ref/libstochqn_ref_{f64,f32}.so`` is the UNMODIFIED reference core
(/root/reference/src/stochqn.c) compiled by ``oracle/build_ref.py``.  This module
loads it and wraps the three optimizers in the same small "stepper" interface that
``oracle/stochqn_np.py`` (NumPy restatement) and the CUDA library use in the tests,
so that one driver (``oracle/driver.py``) can record comparable traces from each.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from stochqn_b200._abi import StochqnABI

_HERE = os.path.dirname(os.path.abspath(__file__))
_CACHE = {}


def ref_path(dtype) -> str:
    tag = "f64" if np.dtype(dtype) == np.float64 else "f32"
    return os.path.join(_HERE, "_ref", "libstochqn_ref_%s.so" % tag)


def have_ref(dtype=np.float64) -> bool:
    return os.path.exists(ref_path(dtype))


def load_ref(dtype=np.float64) -> StochqnABI:
    key = np.dtype(dtype).name
    if key not in _CACHE:
        path = ref_path(dtype)
        if not os.path.exists(path):
            from . import build_ref
            build_ref.build()
        real = C.c_double if np.dtype(dtype) == np.float64 else C.c_float
        _CACHE[key] = StochqnABI(C.CDLL(path), real)
    return _CACHE[key]


def _addr(a):
    return None if a is None else a.ctypes.data


class _RefBase:
    """Shared plumbing: out-params, request labelling, numpy views of workspace arrays."""

    def __init__(self, dtype):
        self.dtype = np.dtype(dtype).type
        self.abi = load_ref(dtype)
        self._req = C.c_void_p()
        self._req_vec = C.c_void_p()
        self._task = C.c_int()
        self._info = C.c_int()
        self.req = None
        self.req_vec = None
        self.req_label = None

    def _view(self, ptr, count):
        addr = C.cast(ptr, C.c_void_p).value
        if not addr:
            return None
        buf = (self.abi.real * count).from_address(addr)
        return np.frombuffer(buf, dtype=self.dtype)

    def _label_req(self, x):
        p = self._req.value
        w = self.ws.contents
        n = w.n
        if p == x.ctypes.data:
            self.req, self.req_label = x, "x"
            return
        for name, lab in (("x_sum", "x_avg"), ("x_avg_prev", "x_avg_prev")):
            if hasattr(w, name):
                a = C.cast(getattr(w, name), C.c_void_p).value
                if a and a == p:
                    self.req, self.req_label = self._view(getattr(w, name), n), lab
                    return
        self.req, self.req_label = None, "?"

    # counters exactly as the wrappers read them back (Rwrapper.c:117-123, pywrapper.pxi:170-207)
    @property
    def niter(self):
        return int(self.ws.contents.niter)

    @property
    def section(self):
        return int(self.ws.contents.section)

    @property
    def bfgs_memory(self):
        return self.ws.contents.bfgs_memory.contents

    def s_slot(self, i):
        m = self.bfgs_memory
        n = self.ws.contents.n
        return self._view(m.s_mem, m.mem_size * n).reshape(m.mem_size, n)[i]

    def y_slot(self, i):
        m = self.bfgs_memory
        n = self.ws.contents.n
        return self._view(m.y_mem, m.mem_size * n).reshape(m.mem_size, n)[i]


class RefOLBFGS(_RefBase):
    kind = "oLBFGS"

    def __init__(self, n, mem_size=10, hess_init=0.0, y_reg=0.0, min_curvature=0.0, check_nan=1, nthreads=1,
                 dtype=np.float64):
        super().__init__(dtype)
        self.ws = self.abi.lib.initialize_oLBFGS(n, mem_size, hess_init, y_reg, min_curvature, check_nan, nthreads)
        assert self.ws, "initialize_oLBFGS returned NULL"

    def run(self, step_size, x, grad):
        ret = self.abi.lib.run_oLBFGS(step_size, _addr(x), _addr(grad), C.byref(self._req), C.byref(self._task),
                                      self.ws, C.byref(self._info))
        self._label_req(x)
        return ret, self._task.value, self._info.value

    def __del__(self):
        if getattr(self, "ws", None):
            self.abi.lib.dealloc_oLBFGS(self.ws)
            self.ws = None


class RefSQN(_RefBase):
    kind = "SQN"

    def __init__(self, n, mem_size=10, bfgs_upd_freq=10, min_curvature=1e-4, use_grad_diff=0, y_reg=0.0,
                 check_nan=1, nthreads=1, dtype=np.float64):
        super().__init__(dtype)
        self.ws = self.abi.lib.initialize_SQN(n, mem_size, bfgs_upd_freq, min_curvature, use_grad_diff, y_reg,
                                              check_nan, nthreads)
        assert self.ws, "initialize_SQN returned NULL"

    def run(self, step_size, x, grad, hess_vec=None):
        ret = self.abi.lib.run_SQN(step_size, _addr(x), _addr(grad), _addr(hess_vec), C.byref(self._req),
                                   C.byref(self._req_vec), C.byref(self._task), self.ws, C.byref(self._info))
        self._label_req(x)
        if self._task.value == 104:
            m = self.bfgs_memory
            self.req_vec = self.s_slot(m.mem_st_ix)
            assert self.req_vec.ctypes.data == self._req_vec.value
        return ret, self._task.value, self._info.value

    def __del__(self):
        if getattr(self, "ws", None):
            self.abi.lib.dealloc_SQN(self.ws)
            self.ws = None


class RefAdaQN(_RefBase):
    kind = "adaQN"

    def __init__(self, n, mem_size=10, fisher_size=100, bfgs_upd_freq=10, max_incr=1.01, min_curvature=1e-4,
                 scal_reg=1e-4, rmsprop_weight=0.9, use_grad_diff=0, y_reg=0.0, check_nan=1, nthreads=1,
                 dtype=np.float64):
        super().__init__(dtype)
        self.ws = self.abi.lib.initialize_adaQN(n, mem_size, fisher_size, bfgs_upd_freq, max_incr, min_curvature,
                                                scal_reg, rmsprop_weight, use_grad_diff, y_reg, check_nan, nthreads)
        assert self.ws, "initialize_adaQN returned NULL"

    @property
    def fisher_memory(self):
        p = self.ws.contents.fisher_memory
        return p.contents if p else None

    @property
    def f_prev(self):
        return float(self.ws.contents.f_prev)

    def run(self, step_size, x, f, grad):
        ret = self.abi.lib.run_adaQN(step_size, _addr(x), f, _addr(grad), C.byref(self._req), C.byref(self._task),
                                     self.ws, C.byref(self._info))
        self._label_req(x)
        return ret, self._task.value, self._info.value

    def __del__(self):
        if getattr(self, "ws", None):
            self.abi.lib.dealloc_adaQN(self.ws)
            self.ws = None
