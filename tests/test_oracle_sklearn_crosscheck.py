"""The scikit-learn-flavoured restatements (oracle/logistic_sk_np.py, oracle/multinomial_np.py) against the installed
scikit-learn's own implementation of the same losses.

The private functions the reference calls (stochqn/_logistic.py:3-4) were removed after scikit-learn 1.0; their
successor `sklearn.linear_model._linear_loss.LinearModelLoss` computes the same loss / gradient / Hessian product with
the sample weights normalised by their sum (and labels {0,1} / class indices instead of {-1,+1} / one-hot).  With
weights that sum to one - which is what the reference's StochasticLogisticRegression passes
(stochqn/_logistic.py:167) - the two must agree to rounding.  This pins the restatements to an independent
implementation; it is not a golden vector of the reference itself (there is none for this path)."""
import numpy as np
import pytest

from oracle import logistic_sk_np as LS
from oracle import multinomial_np as MN

sk = pytest.importorskip("sklearn.linear_model._linear_loss")
from sklearn._loss.loss import HalfBinomialLoss, HalfMultinomialLoss  # noqa: E402


@pytest.mark.parametrize("fit_intercept", [False, True])
def test_binary_restatement_matches_sklearn_linear_model_loss(fit_intercept):
    rng = np.random.default_rng(0)
    n, d, alpha = 300, 9, 0.37
    X = rng.standard_normal((n, d))
    y01 = (rng.random(n) < 0.4).astype(np.float64)
    sw = 0.2 + rng.random(n)
    sw /= sw.sum()
    w = rng.standard_normal(d + fit_intercept)
    s = rng.standard_normal(d + fit_intercept)
    lml = sk.LinearModelLoss(base_loss=HalfBinomialLoss(), fit_intercept=fit_intercept)
    loss_sk, grad_sk = lml.loss_gradient(w, X, y01, sample_weight=sw, l2_reg_strength=alpha)
    _, hessp = lml.gradient_hessian_product(w, X, y01, sample_weight=sw, l2_reg_strength=alpha)
    loss, grad = LS.logistic_loss_and_grad(w, X, 2.0 * y01 - 1.0, alpha, sw)
    assert abs(loss - loss_sk) <= 1e-12 * max(1.0, abs(loss_sk))
    assert np.max(np.abs(grad - grad_sk)) <= 1e-13
    assert np.max(np.abs(LS.logistic_hess_vec(w, s, X, 2.0 * y01 - 1.0, alpha, sw) - hessp(s))) <= 1e-13


@pytest.mark.parametrize("fit_intercept", [False, True])
def test_multinomial_restatement_matches_sklearn_linear_model_loss(fit_intercept):
    rng = np.random.default_rng(1)
    n, d, K, alpha = 240, 7, 5, 0.21
    X = rng.standard_normal((n, d))
    labels = rng.integers(0, K, n)
    Y = np.eye(K)[labels]
    sw = 0.2 + rng.random(n)
    sw /= sw.sum()
    W = rng.standard_normal((K, d + fit_intercept))
    V = rng.standard_normal((K, d + fit_intercept))
    lml = sk.LinearModelLoss(base_loss=HalfMultinomialLoss(n_classes=K), fit_intercept=fit_intercept)
    loss_sk, grad_sk = lml.loss_gradient(W, X, labels.astype(np.float64), sample_weight=sw, l2_reg_strength=alpha)
    _, hessp = lml.gradient_hessian_product(W, X, labels.astype(np.float64), sample_weight=sw, l2_reg_strength=alpha)
    loss, grad, _ = MN.multinomial_loss_grad(W.ravel(), X, Y, alpha, sw)
    assert abs(loss - loss_sk) <= 1e-12 * max(1.0, abs(loss_sk))
    assert np.max(np.abs(grad.reshape(K, -1) - grad_sk)) <= 1e-13
    hv = MN.multinomial_hess_vec(W.ravel(), V.ravel(), X, Y, alpha, sw)
    assert np.max(np.abs(hv.reshape(K, -1) - hessp(V))) <= 1e-13
