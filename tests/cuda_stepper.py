"""Adapter that lets oracle/driver.run_trace drive the CUDA library through its C ABI.

mode="device": x / grad / hess_vec are torch CUDA tensors, the library gets raw device
pointers and `*req` comes back as a device pointer (native mode).
mode="host": x / grad / hess_vec are NumPy arrays, the library stages them itself and `*req`
points at a host mirror (drop-in compatibility mode, the way example/c_rosen.c calls it).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from stochqn_b200 import _lib


class _RawDeviceArray:
    """Minimal __cuda_array_interface__ carrier for a pointer handed out by the C ABI."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class CudaStepper:
    def __init__(self, kind, x0, dtype=np.float64, mode="device", **kw):
        import torch

        self.torch = torch
        self.kind = kind
        self.mode = mode
        self.dtype = np.dtype(dtype).type
        self.abi = _lib.load(dtype)
        lib = self.abi.lib
        n = len(x0)
        self.n = n
        tdt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
        if mode == "device":
            off = 1 if kw.get("misalign") else 0      # views that start one element into an allocation: not 16-byte aligned
            self.x = torch.zeros(n + off, device="cuda", dtype=tdt)[off:]
            self.x.copy_(torch.tensor(np.asarray(x0, dtype=dtype), device="cuda", dtype=tdt))
            self.grad = torch.zeros(n + off, device="cuda", dtype=tdt)[off:]
            self.hess_vec = torch.zeros(n + off, device="cuda", dtype=tdt)[off:]
        else:
            self.x = np.array(x0, dtype=dtype)
            self.grad = np.zeros(n, dtype)
            self.hess_vec = np.zeros(n, dtype)
        if kind == "oLBFGS":
            self.ws = lib.initialize_oLBFGS(n, kw.get("mem_size", 10), kw.get("hess_init", 0.0), kw.get("y_reg", 0.0),
                                            kw.get("min_curvature", 0.0), kw.get("check_nan", 1), 1)
        elif kind == "SQN":
            self.ws = lib.initialize_SQN(n, kw.get("mem_size", 10), kw.get("bfgs_upd_freq", 10),
                                         kw.get("min_curvature", 1e-4), kw.get("use_grad_diff", 0), kw.get("y_reg", 0.0),
                                         kw.get("check_nan", 1), 1)
        else:
            self.ws = lib.initialize_adaQN(n, kw.get("mem_size", 10), kw.get("fisher_size", 100),
                                           kw.get("bfgs_upd_freq", 10), kw.get("max_incr", 1.01),
                                           kw.get("min_curvature", 1e-4), kw.get("scal_reg", 1e-4),
                                           kw.get("rmsprop_weight", 0.9), kw.get("use_grad_diff", 0),
                                           kw.get("y_reg", 0.0), kw.get("check_nan", 1), 1)
        if not self.ws:
            raise RuntimeError("initialize_%s failed: %s" % (kind, _lib.last_error(self.abi)))
        if "one_launch_max_n" in kw:      # 0: force the K1 -> K2 -> K3 route; large: force the fused one-launch step
            assert lib.stochqn_b200_set_option(self.ws, _lib.OPT_ONE_LAUNCH_MAX_N, int(kw["one_launch_max_n"])) == 0
        if "grad_writeback" in kw:
            lib.stochqn_b200_set_option(self.ws, _lib.OPT_GRAD_WRITEBACK, int(kw["grad_writeback"]))
        self._req = C.c_void_p()
        self._req_vec = C.c_void_p()
        self._task = C.c_int()
        self._info = C.c_int()
        self.req_label = None

    # -- raw pointers -------------------------------------------------------------------------
    def _p(self, a):
        return a.data_ptr() if self.mode == "device" else a.ctypes.data

    def call(self, step_size, f=0.0):
        lib = self.abi.lib
        if self.kind == "oLBFGS":
            ret = lib.run_oLBFGS(step_size, self._p(self.x), self._p(self.grad), C.byref(self._req), C.byref(self._task),
                                 self.ws, C.byref(self._info))
        elif self.kind == "SQN":
            ret = lib.run_SQN(step_size, self._p(self.x), self._p(self.grad), self._p(self.hess_vec), C.byref(self._req),
                              C.byref(self._req_vec), C.byref(self._task), self.ws, C.byref(self._info))
        else:
            ret = lib.run_adaQN(step_size, self._p(self.x), f, self._p(self.grad), C.byref(self._req),
                                C.byref(self._task), self.ws, C.byref(self._info))
        w = self.ws.contents
        p = self._req.value
        if p == self._p(self.x):
            self.req_label = "x"
        elif self.mode == "device":
            lab = "?"
            for name, l in (("x_sum", "x_avg"), ("x_avg_prev", "x_avg_prev")):
                if hasattr(w, name) and C.cast(getattr(w, name), C.c_void_p).value == p:
                    lab = l
            self.req_label = lab
        else:
            # host mirrors carry no identity; the section tells which buffer was published
            sec = w.section
            if self.kind == "SQN":
                self.req_label = {2: "x_avg_prev", 3: "x_avg", 4: "x_avg"}.get(sec, "?")
            else:
                self.req_label = {2: "x_avg_prev", 3: "x_avg_prev", 4: "x_avg", 5: "x_avg"}.get(sec, "?")
        return ret, self._task.value, self._info.value

    def _dev_to_np(self, ptr):
        """float64 NumPy copy of the n-vector at raw device address `ptr` (zero-copy torch view, then D2H)."""
        typestr = "<f8" if self.dtype is np.float64 else "<f4"
        raw = _RawDeviceArray(ptr, self.n, typestr)
        t = self.torch.as_tensor(raw, device="cuda")
        return t.cpu().numpy().astype(np.float64)

    def _host_view(self, ptr):
        buf = (self.abi.real * self.n).from_address(ptr)
        return np.frombuffer(buf, dtype=self.dtype).astype(np.float64)

    def read(self, name):
        if name in ("x", "grad"):
            a = getattr(self, name)
            return a.cpu().numpy().astype(np.float64) if self.mode == "device" else a.astype(np.float64)
        ptr = self._req.value if name == "req" else self._req_vec.value
        return self._dev_to_np(ptr) if self.mode == "device" else self._host_view(ptr)

    def write(self, name, arr):
        a = getattr(self, name)
        if self.mode == "device":
            a.copy_(self.torch.from_numpy(np.asarray(arr, dtype=self.dtype)))
        else:
            a[:] = arr

    def counters(self):
        w = self.ws.contents
        m = w.bfgs_memory.contents
        c = dict(niter=int(w.niter), section=int(w.section), mem_used=int(m.mem_used), mem_st_ix=int(m.mem_st_ix))
        if self.kind == "adaQN":
            fm = w.fisher_memory
            c["fisher_used"] = int(fm.contents.mem_used) if fm else 0
            c["fisher_st_ix"] = int(fm.contents.mem_st_ix) if fm else 0
            c["f_prev"] = float(w.f_prev)
        return c

    def slot(self, which, i):
        """Row i of s_mem / y_mem as float64 NumPy (device memory read back)."""
        m = self.ws.contents.bfgs_memory.contents
        ld = self.abi.lib.stochqn_b200_row_stride(self.ws)
        base = C.cast(m.s_mem if which == "s" else m.y_mem, C.c_void_p).value
        esz = C.sizeof(self.abi.real)
        return self._dev_to_np(base + i * ld * esz)

    def one_launch_steps(self):
        return int(_lib.get_stat(self.abi, self.ws, _lib.STAT_ONE_LAUNCH_STEPS))

    def close(self):
        if self.ws:
            lib = self.abi.lib
            {"oLBFGS": lib.dealloc_oLBFGS, "SQN": lib.dealloc_SQN, "adaQN": lib.dealloc_adaQN}[self.kind](self.ws)
            self.ws = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
