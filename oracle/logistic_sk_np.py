"""TEST INFRASTRUCTURE ONLY - NumPy restatement of the two-class logistic arithmetic the reference's Python layer
delegates to scikit-learn (call sites: /root/reference/stochqn/_logistic.py:23-30).

The arithmetic is third-party: private functions `_logistic_loss_and_grad` and `_logistic_grad_hess` of
`sklearn/linear_model/_logistic.py`, present only in scikit-learn <= 1.0 (last release carrying them: 1.0.2; the
reference pins no version).  Not vendored under /root/reference and not importable here (scikit-learn 1.9 removed
them), and the reference holds no test or golden vector for this path.  This file restates the published algorithm of
scikit-learn 1.0.2; `tests/test_oracle_sklearn_crosscheck.py` cross-checks it (and oracle/multinomial_np.py) against
the successor of those functions in the installed scikit-learn (`LinearModelLoss` with `HalfBinomialLoss` /
`HalfMultinomialLoss`), which computes the same loss with the sample weights normalised to sum 1:

    y in {-1,+1};  w has n_features (+1: intercept LAST) entries;  z = X w[:d] + c;  q = sigmoid(y z)
    loss     = -sum(sw * log q) + alpha/2 * |w[:d]|^2
    grad     = [ X'z0 + alpha w[:d] ; sum(z0) ]              z0 = sw (q - 1) y
    hessp(s) = [ X'(d (X s[:d] + s[-1])) + alpha s[:d] ; sum(d (X s[:d] + s[-1])) ]      d = sw q (1 - q)
Sums over samples, not means.
"""
from __future__ import annotations

import numpy as np


def _parts(w, X):
    w = np.asarray(w, np.float64)
    d = X.shape[1]
    if w.size == d + 1:
        return w[:d], float(w[d]), True
    return w, 0.0, False


def logistic_loss_and_grad(w, X, y, alpha, sample_weight=None):
    X = np.asarray(X, np.float64)
    y = np.asarray(y, np.float64)
    sw = np.ones(X.shape[0]) if sample_weight is None else np.asarray(sample_weight, np.float64)
    wd, c, icpt = _parts(w, X)
    yz = y * (X @ wd + c)
    loss = float(np.sum(sw * np.logaddexp(0.0, -yz)) + 0.5 * alpha * (wd @ wd))
    q = 1.0 / (1.0 + np.exp(-yz))
    z0 = sw * (q - 1.0) * y
    g = np.empty(wd.size + (1 if icpt else 0))
    g[:wd.size] = X.T @ z0 + alpha * wd
    if icpt:
        g[-1] = z0.sum()
    return loss, g


def logistic_hess_vec(w, s, X, y, alpha, sample_weight=None):
    X = np.asarray(X, np.float64)
    y = np.asarray(y, np.float64)
    s = np.asarray(s, np.float64)
    sw = np.ones(X.shape[0]) if sample_weight is None else np.asarray(sample_weight, np.float64)
    wd, c, icpt = _parts(w, X)
    q = 1.0 / (1.0 + np.exp(-y * (X @ wd + c)))
    dd = sw * q * (1.0 - q)
    t = X @ s[:wd.size] + (s[-1] if icpt else 0.0)
    out = np.empty_like(s)
    out[:wd.size] = X.T @ (dd * t) + alpha * s[:wd.size]
    if icpt:
        out[-1] = np.sum(dd * t)
    return out
