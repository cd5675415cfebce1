"""TEST INFRASTRUCTURE ONLY - NumPy restatement of the multinomial-logistic arithmetic the reference delegates to
scikit-learn (call sites: /root/reference/stochqn/_logistic.py:7-13, README.md:109,225, notebook cell 3).

PARITY UNPINNED.  The arithmetic is third-party: private functions `_multinomial_loss`, `_multinomial_loss_grad`,
`_multinomial_grad_hess` of `sklearn/linear_model/_logistic.py`, which exist only in scikit-learn <= 1.0 (last
release carrying them: 1.0.2; the reference pins no version: requirements.txt:3, setup.py:45).  That package is
neither vendored under /root/reference nor importable here (scikit-learn 1.9 removed them), and the reference holds no
test or golden vector for this path, so this file restates the published algorithm of scikit-learn 1.0.2 and nothing
checks it against the reference's own output.  It IS cross-checked against an independent implementation: the installed
scikit-learn's `LinearModelLoss(HalfMultinomialLoss)`, successor of those functions, agrees to 1e-13 on loss, gradient
and Hessian-vector product when the sample weights sum to one (tests/test_oracle_sklearn_crosscheck.py):

    w            reshaped (n_classes, n_features [+1]) row-major, intercept = last column
    p            = X @ w.T + intercept ;  p -= logsumexp(p, axis=1) ;  loss = -(sw * Y * p).sum() + 0.5*alpha*||w||^2 ; p = exp(p)
    grad[:, :d]  = (sw * (p - Y)).T @ X + alpha * w ;  grad[:, -1] = (sw * (p - Y)).sum(axis=0)
    hessp(v)     : r = X @ v.T + v_intercept ; r += (-p * r).sum(axis=1) ; r *= p ; r *= sw ;
                   out[:, :d] = r.T @ X + alpha * v ; out[:, -1] = r.sum(axis=0)
Sums over samples, not means (the reference's Python layer leaves any averaging to the caller).
"""
from __future__ import annotations

import numpy as np


def _split(w, n_classes, n_features):
    w = np.asarray(w, np.float64).reshape(n_classes, -1)
    fit_intercept = w.shape[1] == n_features + 1
    if fit_intercept:
        return w[:, :-1], w[:, -1], True
    return w, 0.0, False


def multinomial_loss(w, X, Y, alpha, sample_weight=None):
    X = np.asarray(X, np.float64)
    Y = np.asarray(Y, np.float64)
    n_classes, n_features = Y.shape[1], X.shape[1]
    sw = np.ones(X.shape[0]) if sample_weight is None else np.asarray(sample_weight, np.float64)
    W, b, _ = _split(w, n_classes, n_features)
    p = X @ W.T + b
    m = p.max(axis=1, keepdims=True)
    lse = m + np.log(np.exp(p - m).sum(axis=1, keepdims=True))
    p = p - lse
    loss = -(sw[:, None] * Y * p).sum() + 0.5 * alpha * float((W * W).sum())
    return loss, np.exp(p), W


def multinomial_loss_grad(w, X, Y, alpha, sample_weight=None):
    X = np.asarray(X, np.float64)
    Y = np.asarray(Y, np.float64)
    n_classes, n_features = Y.shape[1], X.shape[1]
    sw = np.ones(X.shape[0]) if sample_weight is None else np.asarray(sample_weight, np.float64)
    loss, p, W = multinomial_loss(w, X, Y, alpha, sw)
    fit_intercept = np.asarray(w).size == n_classes * (n_features + 1)
    grad = np.zeros((n_classes, n_features + int(fit_intercept)))
    diff = sw[:, None] * (p - Y)
    grad[:, :n_features] = diff.T @ X + alpha * W
    if fit_intercept:
        grad[:, -1] = diff.sum(axis=0)
    return loss, grad.ravel(), p


def multinomial_hess_vec(w, v, X, Y, alpha, sample_weight=None):
    X = np.asarray(X, np.float64)
    Y = np.asarray(Y, np.float64)
    n_classes, n_features = Y.shape[1], X.shape[1]
    sw = np.ones(X.shape[0]) if sample_weight is None else np.asarray(sample_weight, np.float64)
    _, _, p = multinomial_loss_grad(w, X, Y, alpha, sw)
    V, vb, fit_intercept = _split(v, n_classes, n_features)
    r = X @ V.T + vb
    r = r + (-p * r).sum(axis=1)[:, None]
    r = r * p
    r = r * sw[:, None]
    out = np.zeros((n_classes, n_features + int(fit_intercept)))
    out[:, :n_features] = r.T @ X + alpha * V
    if fit_intercept:
        out[:, -1] = r.sum(axis=0)
    return out.ravel()
