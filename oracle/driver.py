"""TEST INFRASTRUCTURE ONLY - one request-loop driver for every implementation of the ABI.

``run_trace`` plays the caller's side of the reference's free-mode protocol
(include/stochqn.h:293-383; loop shape of example/c_rosen.c:99-118): call, look at
``task``, evaluate what was asked at ``*req`` (and ``*req_vec``), hand it back, call
again - and records after every call the tuple the reference's wrappers read back
(src/Rwrapper.c:117-123,149-156,185-194): task, return value, niter, section, mem_used,
mem_st_ix, info (+ Fisher counters, f_prev).  The same driver runs

    * the NumPy restatement      (oracle/stochqn_np.py)     via HostStepper
    * the reference C library    (oracle/ref_lib.py)        via HostStepper
    * the CUDA library           (tests/cuda_stepper.py)    via its own adapter

so traces are comparable field by field.
"""
from __future__ import annotations

import numpy as np

CALC_GRAD, CALC_GRAD_SAME_BATCH, CALC_GRAD_BIG_BATCH, CALC_HESS_VEC, CALC_FUN_VAL_BATCH = 101, 102, 103, 104, 105


class HostStepper:
    """Adapter for optimizers that work on host NumPy arrays in place (oracle / reference)."""

    def __init__(self, opt, x0):
        self.opt = opt
        self.kind = opt.kind
        dt = opt.dtype
        self.x = np.array(x0, dtype=dt)
        self.grad = np.zeros_like(self.x)
        self.hess_vec = np.zeros_like(self.x)

    def call(self, step_size, f=0.0):
        if self.kind == "oLBFGS":
            return self.opt.run(step_size, self.x, self.grad)
        if self.kind == "SQN":
            return self.opt.run(step_size, self.x, self.grad, self.hess_vec)
        return self.opt.run(step_size, self.x, f, self.grad)

    @property
    def req_label(self):
        return self.opt.req_label

    def read(self, name):
        if name == "x":
            return self.x.astype(np.float64)
        if name == "req":
            return np.asarray(self.opt.req, dtype=np.float64).copy()
        if name == "req_vec":
            return np.asarray(self.opt.req_vec, dtype=np.float64).copy()
        if name == "grad":
            return self.grad.astype(np.float64)
        raise KeyError(name)

    def write(self, name, arr):
        getattr(self, name)[:] = arr

    def counters(self):
        o = self.opt
        m = o.bfgs_memory
        c = dict(niter=int(o.niter), section=int(o.section), mem_used=int(m.mem_used), mem_st_ix=int(m.mem_st_ix))
        if self.kind == "adaQN":
            fm = o.fisher_memory
            c["fisher_used"] = int(fm.mem_used) if fm is not None else 0
            c["fisher_st_ix"] = int(fm.mem_st_ix) if fm is not None else 0
            c["f_prev"] = float(o.f_prev)
        return c


def run_trace(stepper, problem, n_calls, step_size, hooks=None, keep_x=False):
    """Run `n_calls` calls of the request loop.  Returns a list of per-call records.

    hooks: optional {call_index: fn(stepper, task, payload_dict)} applied just before the
    call to tamper with what is handed back (forced y = 0, NaN gradient, huge f ...).
    """
    trace = []
    ret, task, info = stepper.call(step_size, 0.0)
    rec = dict(task=int(task), ret=int(ret), info=int(info), req=stepper.req_label, **stepper.counters())
    xs = stepper.read("x")
    rec["x_norm"] = float(np.linalg.norm(xs))
    if keep_x:
        rec["x"] = xs
        rec["grad"] = stepper.read("grad")
    trace.append(rec)
    for c in range(1, n_calls):
        f = 0.0
        payload = {}
        if task in (CALC_GRAD, CALC_GRAD_SAME_BATCH, CALC_GRAD_BIG_BATCH):
            kind = {CALC_GRAD: "new", CALC_GRAD_SAME_BATCH: "same", CALC_GRAD_BIG_BATCH: "big"}[task]
            payload["grad"] = problem.grad(stepper.read("req"), kind)
        elif task == CALC_HESS_VEC:
            payload["hess_vec"] = problem.hess_vec(stepper.read("req"), stepper.read("req_vec"))
        elif task == CALC_FUN_VAL_BATCH:
            payload["f"] = problem.fun(stepper.read("req"))
        else:
            raise RuntimeError("optimizer returned task %r" % (task,))
        if hooks and c in hooks:
            hooks[c](stepper, task, payload)
        if "grad" in payload:
            stepper.write("grad", payload["grad"])
        if "hess_vec" in payload:
            stepper.write("hess_vec", payload["hess_vec"])
        f = payload.get("f", 0.0)
        ret, task, info = stepper.call(step_size, f)
        rec = dict(task=int(task), ret=int(ret), info=int(info), req=stepper.req_label, **stepper.counters())
        xs = stepper.read("x")
        rec["x_norm"] = float(np.linalg.norm(xs))
        if keep_x:
            rec["x"] = xs
            rec["grad"] = stepper.read("grad")
        trace.append(rec)
    return trace


DISCRETE_FIELDS = ("task", "ret", "info", "req", "niter", "section", "mem_used", "mem_st_ix",
                   "fisher_used", "fisher_st_ix")


def discrete(trace):
    """The bit-exact part of a trace: everything that is an integer or a label."""
    return [tuple(r.get(k) for k in DISCRETE_FIELDS) for r in trace]
