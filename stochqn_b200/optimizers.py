"""Free-mode optimizer objects over the CUDA C ABI.

Host-side mirror of the reference's Python free-mode classes
(stochqn/_optimizers.py:882-1364: ``oLBFGS_free``, ``SQN_free``, ``adaQN_free`` with
``run_optimizer`` / ``update_gradient`` / ``update_hess_vec`` / ``update_function``): same
class names, constructor arguments, defaults, validation and request dictionaries.  What
changes is where the arrays live:

* ``x`` may be a **torch CUDA tensor** (native mode): the optimizer state, ``x`` and the
  gradient buffer stay in HBM, ``requested_on`` comes back as zero-copy CUDA tensor views,
  and ``update_gradient`` accepts CUDA tensors (a device-to-device copy) or writes can go
  straight into ``self.gradient``.
* ``x`` may be a **NumPy array** (compatibility mode, exactly the reference's calling
  convention): the library stages the arrays through the GPU itself.

Unlike the reference (whose holders use ``np.empty``: uninitialised backup buffers, the
"first run is not reproducible" known issue of its README), every buffer starts zeroed.
There is no CPU implementation behind these classes: they fail loudly without the CUDA library.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._abi import INFO_NAMES, TASK_NAMES

task_dct = dict(TASK_NAMES)     # same names as the reference's dictionaries (stochqn/_optimizers.py:15-29)
info_dct = dict(INFO_NAMES)


class _RawDeviceArray:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 2}


MAX_MEM_SIZE = 32       # kMaxMem of csrc/kernels.cuh: initialize_*() returns NULL above it


def _is_torch(a):
    return type(a).__module__.split(".")[0] == "torch"


class _StochQN_free:
    def _take_common_inputs(self, mem_size, min_curvature, y_reg, check_nan, nthreads, use_float):
        assert mem_size > 0
        assert isinstance(mem_size, int)
        if mem_size > MAX_MEM_SIZE:
            raise ValueError("mem_size = %d: the CUDA build keeps at most %d correction pairs (the compact-form solve "
                             "runs in one warp); the reference accepts any size" % (mem_size, MAX_MEM_SIZE))
        if min_curvature is not None:
            assert min_curvature > 0
        else:
            min_curvature = 0
        if y_reg is not None:
            assert y_reg > 0
        else:
            y_reg = 0
        if nthreads is None or nthreads <= 0:
            nthreads = 1          # advisory only: the GPU build has no host threads to configure
        self.mem_size = mem_size
        self.min_curvature = min_curvature
        self.y_reg = y_reg
        self.check_nan = bool(check_nan)
        self.nthreads = int(nthreads)
        self.use_float = bool(use_float)
        self.c_real_t = C.c_float if self.use_float else C.c_double
        self.np_real_t = np.float32 if self.use_float else np.float64
        self.initialized = False
        self._ws = None
        self._abi = None

    # -- buffers ------------------------------------------------------------------------------
    def _alloc_like(self, x, n):
        if _is_torch(x):
            import torch
            return torch.zeros(n, dtype=x.dtype, device=x.device)
        return np.zeros(n, dtype=self.np_real_t)

    def _check_x(self, x):
        if _is_torch(x):
            import torch
            want = torch.float32 if self.use_float else torch.float64
            if x.dtype != want:
                raise ValueError("'x' has wrong dtype.")
            if not x.is_cuda or not x.is_contiguous() or x.dim() != 1:
                raise ValueError("'x' must be a contiguous 1-d CUDA tensor (or a NumPy array).")
        else:
            assert isinstance(x, np.ndarray)
            if x.dtype != self.np_real_t:
                raise ValueError("'x' has wrong dtype.")
            if not x.flags["C_CONTIGUOUS"] or x.ndim != 1:
                raise ValueError("'x' must be a contiguous 1-d array.")

    @staticmethod
    def _ptr(a):
        return a.data_ptr() if _is_torch(a) else a.ctypes.data

    def _wrap_req(self, ptr, x, n):
        """View of the n-vector the library points `*req` at (no copy)."""
        if ptr == self._ptr(x):
            return x
        if _is_torch(x):
            import torch
            return torch.as_tensor(_RawDeviceArray(ptr, n, "<f4" if self.use_float else "<f8"), device=x.device)
        buf = (self.c_real_t * n).from_address(ptr)
        return np.frombuffer(buf, dtype=self.np_real_t)

    def _store(self, dst, src):
        if _is_torch(dst):
            import torch
            if not _is_torch(src):
                src = torch.as_tensor(np.asarray(src, dtype=self.np_real_t).reshape(-1))
            dst.copy_(src.reshape(-1), non_blocking=True)
        else:
            if _is_torch(src):
                src = src.detach().cpu().numpy()
            dst[:] = np.asarray(src, dtype=self.np_real_t).reshape(-1)

    def update_gradient(self, gradient):
        """Pass the requested gradient to the optimizer (evaluated at "requested_on")."""
        self._store(self.gradient, gradient)

    # -- library plumbing ---------------------------------------------------------------------
    def _load(self):
        self._abi = _lib.load(self.np_real_t)
        return self._abi.lib

    def set_stream(self, cuda_stream):
        """Enqueue this optimizer's kernels on `cuda_stream` (an int handle, e.g.
        ``torch.cuda.current_stream().cuda_stream``) instead of the legacy default stream."""
        self._abi.lib.stochqn_b200_set_stream(self._ws, C.c_void_p(cuda_stream))

    def set_comm(self, comm, n_global):
        """Shard this optimizer: `comm` from stochqn_b200.distributed.init_comm, n_global = total length."""
        if self._abi.lib.stochqn_b200_set_comm(self._ws, comm, n_global) != 0:
            raise RuntimeError(_lib.last_error(self._abi))

    def _counters(self):
        w = self._ws.contents
        m = w.bfgs_memory.contents
        return w, m

    @property
    def niter(self):
        return int(self._ws.contents.niter) if self._ws else 0

    def _request(self, ret, task, info, req_arr):
        if ret == -1000:
            raise ValueError("optimizer received invalid input: " + _lib.last_error(self._abi))
        return {
            "task": task_dct[task],
            "requested_on": req_arr,
            "info": {
                "x_changed_in_run": bool(ret),
                "iteration_number": int(self._ws.contents.niter),
                "iteration_info": info_dct[info],
            },
        }

    def _free(self, name):
        if getattr(self, "_ws", None):
            getattr(self._abi.lib, name)(self._ws)
            self._ws = None

    def __getstate__(self):
        raise TypeError("device-resident optimizer state cannot be pickled directly; use stochqn_b200_export / _import")


class oLBFGS_free(_StochQN_free):
    """oLBFGS optimizer (free mode) - reference stochqn/_optimizers.py:929-1045.

    Requests: calc_grad, then calc_grad_same_batch (skipped when a step was rejected)."""

    def __init__(self, mem_size=10, hess_init=None, min_curvature=1e-4, y_reg=None,
                 check_nan=True, nthreads=-1, use_float=False):
        self._take_common_inputs(mem_size, min_curvature, y_reg, check_nan, nthreads, use_float)
        if hess_init is not None:
            assert hess_init > 0
        else:
            hess_init = 0
        self.hess_init = hess_init

    def _initialize(self, x):
        lib = self._load()
        n = x.shape[0]
        self._ws = lib.initialize_oLBFGS(n, self.mem_size, self.hess_init, self.y_reg, self.min_curvature,
                                         int(self.check_nan), self.nthreads)
        if not self._ws:
            raise MemoryError("initialize_oLBFGS failed: " + _lib.last_error(self._abi))
        self.gradient = self._alloc_like(x, n)
        self._req = C.c_void_p()
        self._task = C.c_int()
        self._info = C.c_int()
        self.initialized = True

    def run_optimizer(self, x, step_size):
        """Continue from where the last request left off; returns the next request dictionary
        ({"task", "requested_on", "info"}) - reference stochqn/_optimizers.py:988-1045."""
        self._check_x(x)
        if not self.initialized:
            self._initialize(x)
        ret = self._abi.lib.run_oLBFGS(step_size, self._ptr(x), self._ptr(self.gradient), C.byref(self._req),
                                       C.byref(self._task), self._ws, C.byref(self._info))
        return self._request(ret, self._task.value, self._info.value, self._wrap_req(self._req.value, x, x.shape[0]))

    def __del__(self):
        try:
            self._free("dealloc_oLBFGS")
        except Exception:
            pass


class SQN_free(_StochQN_free):
    """SQN optimizer (free mode) - reference stochqn/_optimizers.py:1048-1189.

    Requests: calc_grad (bfgs_upd_freq times), then calc_hess_vec, or calc_grad_big_batch with
    use_grad_diff.  For calc_hess_vec "requested_on" is a tuple (point, vector)."""

    def __init__(self, mem_size=10, bfgs_upd_freq=20, min_curvature=1e-4, y_reg=None, use_grad_diff=False,
                 check_nan=True, nthreads=-1, use_float=False):
        self._take_common_inputs(mem_size, min_curvature, y_reg, check_nan, nthreads, use_float)
        assert bfgs_upd_freq > 0
        self.bfgs_upd_freq = int(bfgs_upd_freq)
        self.use_grad_diff = bool(use_grad_diff)

    def _initialize(self, x):
        lib = self._load()
        n = x.shape[0]
        self._ws = lib.initialize_SQN(n, self.mem_size, self.bfgs_upd_freq, self.min_curvature,
                                      int(self.use_grad_diff), self.y_reg, int(self.check_nan), self.nthreads)
        if not self._ws:
            raise MemoryError("initialize_SQN failed: " + _lib.last_error(self._abi))
        self.gradient = self._alloc_like(x, n)
        self.hess_vec = self._alloc_like(x, n if not self.use_grad_diff else 1)
        self._req = C.c_void_p()
        self._req_vec = C.c_void_p()
        self._task = C.c_int()
        self._info = C.c_int()
        self.initialized = True

    def update_hess_vec(self, hess_vec):
        """Pass the requested Hessian-vector product (Hessian at requested_on[0] times requested_on[1])."""
        self._store(self.hess_vec, hess_vec)

    def run_optimizer(self, x, step_size):
        self._check_x(x)
        if not self.initialized:
            self._initialize(x)
        n = x.shape[0]
        ret = self._abi.lib.run_SQN(step_size, self._ptr(x), self._ptr(self.gradient), self._ptr(self.hess_vec),
                                    C.byref(self._req), C.byref(self._req_vec), C.byref(self._task), self._ws,
                                    C.byref(self._info))
        req_arr = self._wrap_req(self._req.value, x, n)
        if self._task.value == 104:
            req_arr = (req_arr, self._wrap_req(self._req_vec.value, x, n))
        return self._request(ret, self._task.value, self._info.value, req_arr)

    def __del__(self):
        try:
            self._free("dealloc_SQN")
        except Exception:
            pass


class adaQN_free(_StochQN_free):
    """adaQN optimizer (free mode) - reference stochqn/_optimizers.py:1192-1364.

    Requests: calc_grad (bfgs_upd_freq times), then calc_fun_val_batch when max_incr is set, then
    calc_grad_big_batch with use_grad_diff (the empirical Fisher product is computed internally
    otherwise)."""

    def __init__(self, mem_size=10, fisher_size=100, bfgs_upd_freq=20, max_incr=1.01, min_curvature=1e-4, scal_reg=1e-4,
                 rmsprop_weight=0.9, y_reg=None, use_grad_diff=False, check_nan=True, nthreads=-1, use_float=False):
        self._take_common_inputs(mem_size, min_curvature, y_reg, check_nan, nthreads, use_float)
        assert bfgs_upd_freq > 0
        bfgs_upd_freq = int(bfgs_upd_freq)
        if not use_grad_diff:
            assert fisher_size > 0
            fisher_size = int(fisher_size)
        else:
            fisher_size = 0
        if max_incr is not None:
            assert max_incr > 0
        else:
            max_incr = 0
        assert scal_reg > 0
        if rmsprop_weight is not None:
            assert rmsprop_weight > 0
            assert rmsprop_weight < 1
        else:
            rmsprop_weight = 0
        self.fisher_size = fisher_size
        self.bfgs_upd_freq = bfgs_upd_freq
        self.max_incr = max_incr
        self.scal_reg = scal_reg
        self.rmsprop_weight = rmsprop_weight
        self.use_grad_diff = bool(use_grad_diff)

    def _initialize(self, x):
        lib = self._load()
        n = x.shape[0]
        self._ws = lib.initialize_adaQN(n, self.mem_size, self.fisher_size, self.bfgs_upd_freq, self.max_incr,
                                        self.min_curvature, self.scal_reg, self.rmsprop_weight,
                                        int(self.use_grad_diff), self.y_reg, int(self.check_nan), self.nthreads)
        if not self._ws:
            raise MemoryError("initialize_adaQN failed: " + _lib.last_error(self._abi))
        self.gradient = self._alloc_like(x, n)
        self.f = 0.0
        self._req = C.c_void_p()
        self._task = C.c_int()
        self._info = C.c_int()
        self.initialized = True

    def update_function(self, fun):
        """Pass the requested objective value (evaluated at "requested_on")."""
        self.f = float(fun)

    def run_optimizer(self, x, step_size):
        self._check_x(x)
        if not self.initialized:
            self._initialize(x)
        ret = self._abi.lib.run_adaQN(step_size, self._ptr(x), self.f, self._ptr(self.gradient), C.byref(self._req),
                                      C.byref(self._task), self._ws, C.byref(self._info))
        return self._request(ret, self._task.value, self._info.value, self._wrap_req(self._req.value, x, x.shape[0]))

    def __del__(self):
        try:
            self._free("dealloc_adaQN")
        except Exception:
            pass
