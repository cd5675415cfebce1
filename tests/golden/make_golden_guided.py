"""Generate tests/golden/guided_f64.json by running the REFERENCE's own guided-mode Python classes.

Run in the build container (needs /root/reference and oracle/_ref):

    python tests/golden/make_golden_guided.py

What runs: the unmodified ``/root/reference/stochqn/_optimizers.py`` (classes ``oLBFGS`` / ``SQN`` / ``adaQN``:
``fit``, ``partial_fit``, long-batch bookkeeping, shuffling, validation split, step-size schedules), imported
as module ``stochqn._optimizers`` of a stand-in package so that the package ``__init__`` (which pulls in
scikit-learn private functions that no longer exist) is not executed.  The two compiled
Cython modules it expects (``_wrapper_double`` / ``_wrapper_float``, not buildable here: ``findblas`` is missing)
are replaced by thin ctypes stand-ins that call the same C entry points (run_oLBFGS / run_SQN / run_adaQN) of
the reference C library compiled in ``oracle/_ref`` and return the tuples ``pywrapper.pxi:161-207`` returns.
The workspace is kept C-side (zero-initialised, R semantics) instead of being re-assembled from the holders'
``np.empty`` buffers each call; the arithmetic and the control flow are the reference's.

Recorded per case: the snapshots of x taken by the epoch / iteration callbacks, the final x, niter, the epoch
reached, and the (task, info, ret) triple of every optimizer call.
"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_lib  # noqa: E402

REF_PKG = "/root/reference/stochqn"
CALLS = []          # (task, info, ret) of every optimizer call of the case being run


def _fake_wrapper(dtype):
    mod = types.ModuleType("stochqn._wrapper_" + ("double" if dtype == np.float64 else "float"))

    def py_run_oLBFGS(h, x, grad, step_size):
        if not hasattr(h, "_ref"):
            b = h.BFGS_mem
            h._ref = ref_lib.RefOLBFGS(h.n, mem_size=b.mem_size, hess_init=h.hess_init, y_reg=b.y_reg,
                                       min_curvature=b.min_curvature, check_nan=h.check_nan, nthreads=1, dtype=dtype)
        r = h._ref
        ret, task, info = r.run(step_size, x, grad)
        CALLS.append((task, info, ret))
        m = r.bfgs_memory
        return ret, r.niter, r.section, int(m.mem_used), int(m.mem_st_ix), task, info, r.req

    def py_run_SQN(h, x, step_size, grad, hess_vec):
        if not hasattr(h, "_ref"):
            b = h.BFGS_mem
            h._ref = ref_lib.RefSQN(h.n, mem_size=b.mem_size, bfgs_upd_freq=b.upd_freq, min_curvature=b.min_curvature,
                                    use_grad_diff=h.use_grad_diff, y_reg=b.y_reg, check_nan=h.check_nan, nthreads=1,
                                    dtype=dtype)
        r = h._ref
        hv = hess_vec if hess_vec.shape[0] == x.shape[0] else np.zeros_like(x)
        ret, task, info = r.run(step_size, x, grad, hv)
        CALLS.append((task, info, ret))
        m = r.bfgs_memory
        return (ret, r.niter, r.section, int(m.mem_used), int(m.mem_st_ix), task, info, r.req,
                r.req_vec if task == 104 else None)

    def py_run_adaQN(h, x, grad, step_size, f):
        if not hasattr(h, "_ref"):
            b = h.BFGS_mem
            fs = 0 if h.use_grad_diff else h.Fisher_mem.mem_size
            h._ref = ref_lib.RefAdaQN(h.n, mem_size=b.mem_size, fisher_size=fs, bfgs_upd_freq=b.upd_freq,
                                      max_incr=h.max_incr, min_curvature=b.min_curvature, scal_reg=h.scal_reg,
                                      rmsprop_weight=h.rmsprop_weight, use_grad_diff=h.use_grad_diff, y_reg=b.y_reg,
                                      check_nan=h.check_nan, nthreads=1, dtype=dtype)
        r = h._ref
        ret, task, info = r.run(step_size, x, f, grad)
        CALLS.append((task, info, ret))
        m = r.bfgs_memory
        fm = r.fisher_memory
        return (ret, r.niter, r.section, int(m.mem_used), int(m.mem_st_ix),
                int(fm.mem_used) if fm is not None else 0, int(fm.mem_st_ix) if fm is not None else 0, r.f_prev,
                task, info, r.req)

    mod.py_run_oLBFGS, mod.py_run_SQN, mod.py_run_adaQN = py_run_oLBFGS, py_run_SQN, py_run_adaQN
    return mod


def import_reference_guided():
    pkg = types.ModuleType("stochqn")
    pkg.__path__ = [REF_PKG]
    sys.modules["stochqn"] = pkg
    sys.modules["stochqn._wrapper_double"] = pkg._wrapper_double = _fake_wrapper(np.float64)
    sys.modules["stochqn._wrapper_float"] = pkg._wrapper_float = _fake_wrapper(np.float32)
    import importlib
    return importlib.import_module("stochqn._optimizers")


def main():
    import warnings

    from guided_support import GUIDED_CASES, drive

    ref = import_reference_guided()
    out = {}
    for name, kind, okw, how in GUIDED_CASES:
        CALLS.clear()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            obj, snaps = drive(getattr(ref, kind), name, okw, how)
        out[name] = dict(x=[float(v) for v in obj.x], niter=int(obj.niter), epoch=int(obj.epoch),
                         snaps=[[float(v) for v in s] for s in snaps], calls=[list(map(int, c)) for c in CALLS])
        infos = sorted(set(c[1] for c in CALLS))
        print("%-28s niter %4d epoch %d calls %4d infos %s |x| %.6f" % (name, obj.niter, obj.epoch, len(CALLS), infos,
                                                                       float(np.linalg.norm(obj.x))))
    path = os.path.join(HERE, "guided_f64.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path)


if __name__ == "__main__":
    main()
