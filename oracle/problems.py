"""TEST INFRASTRUCTURE ONLY - deterministic NumPy objectives used to serve optimizer requests.

Each problem answers the five free-mode requests of the reference API
(include/stochqn.h:268-275) from a point handed over as a NumPy array, in float64,
so that the oracle, the reference library and the CUDA library are all fed
bit-identical gradients.  Batches are a pure function of a call counter: no RNG state
is shared with the optimizers.
"""
from __future__ import annotations

import numpy as np


class Quadratic:
    """f(x) = 0.5 x'Ax - b'x with a fixed SPD matrix (the 6-variable case of SURVEY.md section 4)."""

    def __init__(self, n=6, seed=0, cond=30.0):
        rng = np.random.default_rng(seed)
        q, _ = np.linalg.qr(rng.standard_normal((n, n)))
        ev = np.linspace(1.0, cond, n)
        self.A = (q * ev) @ q.T
        self.A = 0.5 * (self.A + self.A.T)
        self.b = rng.standard_normal(n)
        self.n = n

    def x0(self):
        return np.linspace(-1.0, 1.0, self.n)

    def grad(self, x, kind):
        return self.A @ x - self.b

    def hess_vec(self, x, v):
        return self.A @ v

    def fun(self, x):
        return float(0.5 * x @ self.A @ x - self.b @ x)


class Rosenbrock:
    """Chained Rosenbrock with the formulas of the reference's example program
    (example/c_rosen.c:13-59).  `example_quirk=True` reproduces that file's Hessian-vector
    routine literally (its first row multiplies p[0] where the analytic Hessian has p[1],
    c_rosen.c:46) - needed to reproduce the example's printed output."""

    def __init__(self, n, example_quirk=False):
        self.n = n
        self.example_quirk = example_quirk

    def x0(self):
        i = np.arange(self.n, dtype=np.uint64)
        h = (i * np.uint64(2654435761)) & np.uint64(0xFFFFFFFF)
        return 0.95 + 1e-4 * (h % np.uint64(1000)).astype(np.float64)

    def fun(self, x):
        x = np.asarray(x, np.float64)
        d1 = x[1:] - x[:-1] ** 2
        d2 = 1.0 - x[:-1]
        return float(np.sum(100.0 * d1 * d1 + d2 * d2))

    def grad(self, x, kind=None):
        x = np.asarray(x, np.float64)
        g = np.empty_like(x)
        g[0] = -400.0 * x[0] * (x[1] - x[0] * x[0]) - 2.0 * (1.0 - x[0])
        g[-1] = 200.0 * (x[-1] - x[-2] * x[-2])
        xm, xc, xp = x[:-2], x[1:-1], x[2:]
        g[1:-1] = 200.0 * (xc - xm * xm) - 400.0 * (xp - xc * xc) * xc - 2.0 * (1.0 - xc)
        return g

    def hess_vec(self, x, p):
        x = np.asarray(x, np.float64)
        p = np.asarray(p, np.float64)
        out = np.zeros_like(x)
        if self.example_quirk:
            out[0] = (1200 * x[0] * x[0] - 400 * x[1] + 2.0) * p[0] - 400 * x[0] * p[0]
        else:
            out[0] = (1200 * x[0] * x[0] - 400 * x[1] + 2.0) * p[0] - 400 * x[0] * p[1]
        out[-1] = -400.0 * x[-2] * p[-2] + 200.0 * p[-1]
        xm, xc, xp = x[:-2], x[1:-1], x[2:]
        out[1:-1] = -400.0 * xm * p[:-2] + (202 + 1200 * xc * xc - 400 * xp) * p[1:-1] - 400.0 * xc * p[2:]
        return out


class Logistic:
    """Binary logistic regression with the closed forms of the reference's R model
    (R/logistic.R:1-37): loss = mean log-loss + lambda*||w||^2, grad = X'(p-y)/N + 2*lambda*w,
    Hv = X'(p(1-p) * Xv)/N + 2*lambda*v.  Batches are consecutive row blocks; a "big batch"
    is the union of the last `big` ordinary batches, as the guided R driver stacks them
    (R/optimizers_guided.R:26-111)."""

    def __init__(self, nrows=4000, ncols=40, batch=200, big=5, seed=1, lam=1e-5):
        rng = np.random.default_rng(seed)
        self.X = rng.standard_normal((nrows, ncols))
        self.X[:, 0] = 1.0
        w_true = rng.standard_normal(ncols)
        p = 1.0 / (1.0 + np.exp(-self.X @ w_true))
        self.y = (rng.random(nrows) < p).astype(np.float64)
        self.n = ncols
        self.batch = batch
        self.big = big
        self.lam = lam
        self.nb = nrows // batch
        self.ib = -1            # index of the current ordinary batch
        self.Xv = self.X[: 2 * batch]
        self.yv = self.y[: 2 * batch]

    def x0(self):
        return np.zeros(self.n)

    def _rows(self, kind):
        if kind == "new":
            self.ib += 1
        if kind in ("new", "same"):
            b = self.ib % self.nb
            return slice(b * self.batch, (b + 1) * self.batch)
        # big batch: the last `big` ordinary batches (wrapping is avoided by the modulus on whole blocks)
        last = self.ib % self.nb
        first = max(0, last - self.big + 1)
        return slice(first * self.batch, (last + 1) * self.batch)

    def grad(self, x, kind):
        r = self._rows(kind)
        X, y = self.X[r], self.y[r]
        p = 1.0 / (1.0 + np.exp(-(X @ x)))
        return X.T @ (p - y) / X.shape[0] + 2.0 * self.lam * x

    def hess_vec(self, x, v):
        r = self._rows("big")
        X = self.X[r]
        p = 1.0 / (1.0 + np.exp(-(X @ x)))
        return X.T @ (p * (1.0 - p) * (X @ v)) / X.shape[0] + 2.0 * self.lam * v

    def fun(self, x):
        z = self.Xv @ x
        p = 1.0 / (1.0 + np.exp(-z))
        eps = 1e-300
        return float(-np.mean(self.yv * np.log(p + eps) + (1 - self.yv) * np.log(1 - p + eps)) + self.lam * x @ x)
