"""BENCH INFRASTRUCTURE ONLY (cpu_baseline legs of bench.py's `secondary` block) - the reference's own C library
(oracle/_ref, unmodified src/stochqn.c) driven through its free-mode request loop on the host cores, with the callbacks
the reference's R / Python layers would evaluate written in NumPy on top of BLAS:

  logistic   R/logistic.R:12-37     grad = X'(p - y)/N + 2*lambda*w ;  Hv = X'(p(1-p) * Xv)/N + 2*lambda*v
  multinomial  stochqn/_logistic.py:7-13 (scikit-learn <= 1.0 arithmetic): P = softmax(X W' + b),
               grad = (sw * (P - Y))' [X 1] + alpha*[W 0]

One matrix-vector (or matrix-matrix) product per pass over the batch and no temporaries of the batch's size: this is
cheaper than what R does (R/logistic.R materialises X * (p - y)), so the figure flatters the reference.  The data are
the SAME rows the GPU leg used (copied from the device).  Not part of the product; never imported by stochqn_b200/.
"""
from __future__ import annotations

import os
import time

import numpy as np

from oracle import ref_lib as R


def _threads():
    return os.cpu_count() or 1


def best_of_threads(fn, **kw):
    """The reference with the better of {1, all} host threads (BLAS and OpenMP pools both limited): oversubscription
    hurts the small configurations, the large ones want every core."""
    from threadpoolctl import threadpool_limits

    best = None
    counts = kw.pop("thread_counts", None) or sorted({1, _threads()})
    for k in counts:
        with threadpool_limits(limits=k):
            r = fn(nthreads=k, **kw)
        if best is None or r["value"] > best["value"]:
            best = r
    return best


def logistic_reference(kind, X, y, batch, steps, warm, L=10, big=20000, lam=1e-5, step=1e-1, mem=10, nthreads=None, max_s=8.0):
    """oLBFGS (config 1) or SQN with Hessian-vector products (config 2) + binary logistic regression.
    X: host (rows, n) with the intercept column first, y: host (rows,).  Batches cycle over the rows given."""
    n = X.shape[1]
    nth = nthreads or _threads()
    if kind == "oLBFGS":
        opt = R.RefOLBFGS(n, mem_size=mem, hess_init=0.0, y_reg=0.0, min_curvature=1e-4, check_nan=1, nthreads=nth)
    else:
        opt = R.RefSQN(n, mem_size=mem, bfgs_upd_freq=L, min_curvature=1e-4, use_grad_diff=0, y_reg=0.0, check_nan=1, nthreads=nth)
    x = np.zeros(n)
    g = np.zeros(n)
    hv = np.zeros(n)
    nb = X.shape[0] // batch
    state = {"b": 0}

    def grad_rows(r0, cnt, at):
        Xb, yb = X[r0:r0 + cnt], y[r0:r0 + cnt]
        p = 1.0 / (1.0 + np.exp(-(Xb @ at)))
        return Xb.T @ (p - yb) / cnt + 2.0 * lam * at

    def call():
        if kind == "oLBFGS":
            return opt.run(step, x, g)
        return opt.run(step, x, g, hv)

    ret, task, info = call()
    t0, it0, t_begin = None, 0, time.perf_counter()
    while True:
        it = opt.niter
        if t0 is None and it >= warm and task == 101:
            t0, it0 = time.perf_counter(), it
        if t0 is not None and task == 101 and (it >= it0 + steps or (it > it0 and time.perf_counter() - t0 > max_s)):
            steps = it - it0          # (the time cap cut the sample short: the rate is over the steps actually done)
            break
        if t0 is None and task == 101 and time.perf_counter() - t_begin > 3 * max_s:
            warm = it                 # warm-up is taking too long on this host: start timing now, say so in `mem_used`
        b = state["b"]
        if task == 101:
            state["b"] = b = (b + 1) % nb
            g[:] = grad_rows(b * batch, batch, opt.req)
        elif task == 102:
            g[:] = grad_rows(b * batch, batch, opt.req)
        elif task == 104:
            cnt = min(big, (b + 1) * batch)
            r0 = (b + 1) * batch - cnt
            Xb = X[r0:r0 + cnt]
            p = 1.0 / (1.0 + np.exp(-(Xb @ opt.req)))
            hv[:] = Xb.T @ (p * (1.0 - p) * (Xb @ opt.req_vec)) / cnt + 2.0 * lam * opt.req_vec
        else:
            raise RuntimeError("unexpected task %d" % task)
        ret, task, info = call()
    dt = time.perf_counter() - t0
    return {"value": steps / dt, "unit": "steps/s", "cores": nth, "kind": "reference", "mem_used": int(opt.bfgs_memory.mem_used),
            "sample": "reference C library (oracle/_ref) + NumPy/BLAS logistic callbacks on the GPU leg's first %d rows, %d warm-up + %d timed steps, %d threads"
                      % (X.shape[0], warm, steps, nth)}


def multinomial_reference(dtype, X, lab, K, batch, steps, warm, L, fisher, use_grad_diff, max_incr, rms, step, alpha=1e-3, mem=10, nval=740, nthreads=None, max_s=12.0,
                          wsum=False, x0=None, fixed_big=False):
    """adaQN + multinomial logistic regression (configs 3 and 5).  X: host (rows, d), lab: host int labels in [0, K).
    wsum: sample weights of one (summed loss) instead of 1/rows; x0: start point; fixed_big: the big batch is always rows [0, big)."""
    d = X.shape[1]
    n = K * (d + 1)
    nth = nthreads or _threads()
    opt = R.RefAdaQN(n, mem_size=mem, fisher_size=max(fisher, 1), bfgs_upd_freq=L, max_incr=max_incr, min_curvature=1e-4, scal_reg=1e-4,
                     rmsprop_weight=rms, use_grad_diff=use_grad_diff, y_reg=0.0, check_nan=1, nthreads=nth, dtype=dtype)
    x = np.zeros(n, dtype) if x0 is None else np.array(x0, dtype=dtype)
    g = np.zeros(n, dtype)
    nb = X.shape[0] // batch
    big = min(X.shape[0], batch * L)
    nval = min(nval, X.shape[0])
    state = {"b": 0, "f": 0.0}

    def loss_grad(r0, cnt, at, want_loss):
        Xb = X[r0:r0 + cnt]
        W = at.reshape(K, d + 1)
        Z = Xb @ W[:, :d].T + W[:, d]
        Z -= Z.max(axis=1, keepdims=True)
        np.exp(Z, out=Z)
        s = Z.sum(axis=1, keepdims=True)
        w = 1.0 if wsum else 1.0 / cnt
        if want_loss:
            pl = Z[np.arange(cnt), lab[r0:r0 + cnt]] / s[:, 0]
            return float(-w * np.sum(np.log(pl)) + 0.5 * alpha * np.sum(W[:, :d].astype(np.float64) ** 2))
        Z /= s
        Z[np.arange(cnt), lab[r0:r0 + cnt]] -= 1.0
        if not wsum:
            Z *= w
        G = np.empty((K, d + 1), dtype)
        np.matmul(Z.T, Xb, out=G[:, :d])
        G[:, :d] += alpha * W[:, :d]
        G[:, d] = Z.sum(axis=0)
        return G.reshape(-1)

    ret, task, info = opt.run(step, x, state["f"], g)
    t0, it0, t_begin = None, 0, time.perf_counter()
    while True:
        it = opt.niter
        if t0 is None and it >= warm and task == 101:
            t0, it0 = time.perf_counter(), it
        if t0 is not None and task == 101 and (it >= it0 + steps or (it > it0 and time.perf_counter() - t0 > max_s)):
            steps = it - it0          # (the time cap cut the sample short: the rate is over the steps actually done)
            break
        if t0 is None and task == 101 and time.perf_counter() - t_begin > 3 * max_s:
            warm = it                 # warm-up is taking too long on this host: start timing now, say so in `mem_used`
        b = state["b"]
        if task == 101:
            state["b"] = b = (b + 1) % nb
            g[:] = loss_grad(b * batch, batch, opt.req, False)
        elif task == 103:
            r0 = 0 if fixed_big else max(0, (b + 1) * batch - big)
            g[:] = loss_grad(r0, big, opt.req, False)
        elif task == 105:
            state["f"] = loss_grad(0, nval, opt.req, True)
        else:
            raise RuntimeError("unexpected task %d" % task)
        ret, task, info = opt.run(step, x, state["f"], g)
    dt = time.perf_counter() - t0
    return {"value": steps / dt, "unit": "steps/s", "cores": nth, "kind": "reference", "mem_used": int(opt.bfgs_memory.mem_used),
            "sample": "reference C library (oracle/_ref, %s) + NumPy/BLAS multinomial callbacks, %d features x %d classes, batch %d, %d warm-up + %d timed steps, %d threads"
                      % ("f64" if np.dtype(dtype) == np.float64 else "f32", d, K, batch, warm, steps, nth)}
