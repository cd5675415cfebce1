"""StochasticLogisticRegression (stochqn_b200/logistic.py) on the GPU against a host twin: the same guided
classes driven by the oracle optimizer (NumPy) with the NumPy restatements of the scikit-learn callbacks the reference
uses (oracle/logistic_sk_np.py, oracle/multinomial_np.py), same starting point, batches, shuffling and split.
(The reference's own class cannot be imported here: the scikit-learn private functions it needs are gone,
SURVEY.md section 8(c); the guided layer underneath is pinned by tests/test_guided_host.py.)"""
import warnings

import numpy as np
import pytest

from guided_support import ORACLE_FREE
from oracle import logistic_sk_np as LS
from oracle import multinomial_np as MN
from stochqn_b200 import guided

pytestmark = pytest.mark.gpu


def _oracle_backed(kind):
    base = getattr(guided, kind)
    return type(kind, (base,), {"_free_class": staticmethod(lambda: ORACLE_FREE[kind])})


def _host_twin(optimizer, X, y, sw, reg, fit_intercept, random_state, okw):
    mult = y.ndim == 2
    if mult:
        funs = dict(grad_fun=lambda w, X, y, sample_weight=None, reg_param=0: MN.multinomial_loss_grad(w, X, y, reg_param, sample_weight)[1],
                    obj_fun=lambda w, X, y, sample_weight=None, reg_param=0: MN.multinomial_loss_grad(w, X, y, reg_param, sample_weight)[0])
        hv = lambda w, v, X, y, sample_weight=None, reg_param=0: MN.multinomial_hess_vec(w, v, X, y, reg_param, sample_weight)
    else:
        funs = dict(grad_fun=lambda w, X, y, sample_weight=None, reg_param=0: LS.logistic_loss_and_grad(w, X, y, reg_param, sample_weight)[1],
                    obj_fun=lambda w, X, y, sample_weight=None, reg_param=0: LS.logistic_loss_and_grad(w, X, y, reg_param, sample_weight)[0])
        hv = lambda w, v, X, y, sample_weight=None, reg_param=0: LS.logistic_hess_vec(w, v, X, y, reg_param, sample_weight)
    sw = np.ones(X.shape[0]) if sw is None else sw.copy()
    sw = sw / sw.sum()
    np.random.seed(random_state)
    w0 = np.random.normal(size=(X.shape[1] + fit_intercept) * (y.shape[1] if mult else 1))
    kw = dict(okw)
    if optimizer == "SQN":
        kw["hess_vec_fun"] = hv
    opt = _oracle_backed(optimizer)(x0=w0, **funs, **kw)
    opt.fit(X, y, sw, {"reg_param": reg})
    return opt


def _data(mult, seed=0, n=900, d=10, K=4):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, d))
    if mult:
        W = rng.standard_normal((K, d))
        lab = np.argmax(X @ W.T + rng.gumbel(size=(n, K)), axis=1)
        y = np.eye(K)[lab]
    else:
        w = rng.standard_normal(d)
        y = np.where(rng.random(n) < 1 / (1 + np.exp(-X @ w)), 1.0, -1.0)
    sw = 0.5 + rng.random(n)
    return X, y, sw


CASES = [
    ("bin_sqn", False, "SQN", dict(batches_per_epoch=6, nepochs=3, bfgs_upd_freq=4, mem_size=5)),
    ("bin_olbfgs_noicpt", False, "oLBFGS", dict(batches_per_epoch=6, nepochs=3, mem_size=5)),
    ("mult_adaqn_fisher", True, "adaQN", dict(batches_per_epoch=6, nepochs=3, bfgs_upd_freq=3, fisher_size=8, mem_size=5)),
    ("mult_sqn", True, "SQN", dict(batches_per_epoch=6, nepochs=2, bfgs_upd_freq=4, mem_size=5)),
]


@pytest.mark.parametrize("name,mult,optimizer,okw", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("container,native", [("numpy", True), ("cuda", True), ("cuda", False)])
def test_estimator_matches_host_twin(name, mult, optimizer, okw, container, native):
    import torch
    from stochqn_b200.logistic import StochasticLogisticRegression

    X, y, sw = _data(mult)
    fit_intercept = "noicpt" not in name
    reg, step = 1e-3, 1e-1
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        twin = _host_twin(optimizer, X, y, sw, reg, fit_intercept, 1,
                          dict(okw, step_size=step, valset_frac=0.1, verbose=False))
        conv = (lambda a: torch.tensor(a, device="cuda")) if container == "cuda" else (lambda a: a)
        m = StochasticLogisticRegression(reg_param=reg, fit_intercept=fit_intercept, random_state=1, optimizer=optimizer,
                                         step_size=step, valset_frac=0.1, verbose=False, **okw)
        m.native_loop = native      # True: the request loop of a mini-batch runs inside the library (stochqn_b200_fit_batch)
        m.fit(conv(X), conv(y), conv(sw))
    served = m.optimizer.native_batches
    assert (served > 0) == native, "native request loop: %d mini-batches" % served
    assert m.optimizer.niter == twin.niter and m.optimizer.epoch == twin.epoch
    x = m.optimizer.x.cpu().numpy()
    scale = max(1.0, float(np.max(np.abs(twin.x))))
    assert np.max(np.abs(x - twin.x)) / scale <= 1e-8
    # fitted attributes and predictions
    d = X.shape[1]
    if mult:
        Wt = twin.x.reshape(y.shape[1], -1)
        scores = X @ Wt[:, :d].T + (Wt[:, -1] if fit_intercept else 0.0)
        pred_ref = np.argmax(scores, axis=1)
        proba_ref = 1 / (1 + np.exp(-scores))
    else:
        z = X @ twin.x[:d] + (twin.x[-1] if fit_intercept else 0.0)
        pred_ref = (1 / (1 + np.exp(-z)) >= .5).astype("uint8")
        proba_ref = np.c_[1 - 1 / (1 + np.exp(-z)), 1 / (1 + np.exp(-z))]
    pred = m.predict(conv(X))
    proba = m.predict_proba(conv(X))
    if container == "cuda":
        pred, proba = pred.cpu().numpy(), proba.cpu().numpy()
    assert np.mean(pred == pred_ref) >= 0.999
    assert np.max(np.abs(proba - proba_ref)) <= 1e-7
    coef = m.coef_ if container == "numpy" else m.coef_.cpu().numpy()
    assert coef.shape == ((y.shape[1], d) if mult else (d,))


def test_estimator_partial_fit_and_float():
    import torch
    from stochqn_b200.logistic import StochasticLogisticRegression

    X, y, _ = _data(False, seed=3)
    m = StochasticLogisticRegression(reg_param=1e-3, optimizer="adaQN", step_size=5e-2, valset_frac=None, verbose=False,
                                     use_float=True, max_incr=None, fisher_size=10, bfgs_upd_freq=3, mem_size=4)
    Xd, yd = torch.tensor(X, device="cuda"), torch.tensor(y, device="cuda")
    for k in range(9):
        m.partial_fit(Xd[100 * k:100 * (k + 1)], yd[100 * k:100 * (k + 1)])
    assert m.optimizer.native_batches > 0
    assert m.is_fitted and m.optimizer.niter == 9 and m.optimizer.x.dtype == torch.float32
    assert bool(torch.isfinite(m.optimizer.x).all())
    assert m.predict(Xd).shape == (X.shape[0],) and m.predict_proba(Xd).shape == (X.shape[0], 2)


@pytest.mark.parametrize("fmt", ["scipy_csr", "scipy_coo", "torch_csr"])
@pytest.mark.parametrize("mult,optimizer", [(False, "oLBFGS"), (True, "adaQN")])
def test_sparse_model_matrix_is_expanded_on_the_device(fmt, mult, optimizer):
    """The reference's estimator keeps scipy CSR inputs (stochqn/_logistic.py:155).  Here a sparse model matrix is expanded once, on
    the device (stochqn_b200_csr_to_dense), into the dense matrix the bundled kernels stream: same kernels, same batches, so the fit
    is bit-identical to the one on the dense array - also for a BibTeX-like 4 %-dense matrix with empty rows and columns."""
    import scipy.sparse as sp
    import torch
    from stochqn_b200.logistic import StochasticLogisticRegression, _csr_to_dense_on_device

    rng = np.random.default_rng(7)
    n, d, K = 700, 83, 5
    Xd = rng.standard_normal((n, d)) * (rng.random((n, d)) < 0.04)
    Xd[17] = 0.0
    Xd[:, 5] = 0.0
    if mult:
        lab = rng.integers(0, K, n)
        y = np.eye(K)[lab]
    else:
        y = np.where(rng.random(n) < 0.5, 1.0, -1.0)
    Xs = {"scipy_csr": sp.csr_matrix(Xd), "scipy_coo": sp.coo_matrix(Xd), "torch_csr": torch.tensor(Xd, device="cuda").to_sparse_csr()}[fmt]
    dense = _csr_to_dense_on_device(Xs, torch.device("cuda"), torch.float64)
    assert dense.shape == (n, d) and np.array_equal(dense.cpu().numpy(), Xd)
    kw = dict(reg_param=1e-3, random_state=1, optimizer=optimizer, step_size=5e-2, valset_frac=None, verbose=False, batches_per_epoch=7, nepochs=2,
              mem_size=4)
    if optimizer == "adaQN":
        kw.update(bfgs_upd_freq=3, fisher_size=6, max_incr=None)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = StochasticLogisticRegression(**kw).fit(Xs, y)
        b = StochasticLogisticRegression(**kw).fit(Xd, y)
    assert a.optimizer.niter == b.optimizer.niter
    assert torch.equal(a.optimizer.x, b.optimizer.x)
    pa, pb = a.predict(Xs), b.predict(Xd)
    assert np.array_equal(np.asarray(pa.cpu() if hasattr(pa, "cpu") else pa), np.asarray(pb))
