"""GPU: the bundled device callbacks against NumPy restatements of the reference's formulas
(example/c_rosen.c:13-41; R/logistic.R:1-37)."""
import ctypes as C

import numpy as np
import pytest

from oracle.problems import Rosenbrock
from stochqn_b200 import _lib

pytestmark = pytest.mark.gpu


def _t(a, dtype):
    import torch
    return torch.tensor(np.asarray(a, dtype=dtype), device="cuda")


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-13), (np.float32, 2e-6)])
@pytest.mark.parametrize("n", [2, 3, 17, 1000, 1001, 4099])
def test_rosenbrock_x0_grad_fun(dtype, tol, n):
    import torch
    lib = _lib.load(dtype).lib
    p = Rosenbrock(n)
    x = torch.empty(n, device="cuda", dtype=torch.float64 if dtype == np.float64 else torch.float32)
    g = torch.empty_like(x)
    f = torch.zeros(1, device="cuda", dtype=torch.float64)
    assert lib.stochqn_b200_rosenbrock_x0(x.data_ptr(), n, 0, None) == 0
    x0 = p.x0().astype(dtype)
    assert np.array_equal(x.cpu().numpy(), x0)                  # integer hash: bit-exact
    xr = _t(np.random.default_rng(n).standard_normal(n) * 0.3 + 1.0, dtype)
    assert lib.stochqn_b200_rosenbrock_grad(xr.data_ptr(), g.data_ptr(), n, 0, n, None, None) == 0
    assert lib.stochqn_b200_rosenbrock_fun(xr.data_ptr(), n, 0, n, None, f.data_ptr(), None) == 0
    xh = xr.cpu().numpy().astype(np.float64)
    gref, fref = p.grad(xh), p.fun(xh)
    assert np.max(np.abs(g.cpu().numpy() - gref)) <= tol * max(np.max(np.abs(gref)), 1.0) * 10
    assert abs(f.item() - fref) <= 1e-12 * abs(fref)


def test_rosenbrock_shards_with_halo_equal_whole():
    import torch
    lib = _lib.load(np.float64).lib
    n = 1003
    xh = np.random.default_rng(3).standard_normal(n) * 0.3 + 1.0
    gref = Rosenbrock(n).grad(xh)
    fsum = 0.0
    for off, cnt in ((0, 400), (400, 301), (701, 302)):
        x = _t(xh[off:off + cnt], np.float64)
        g = torch.empty_like(x)
        halo = _t([xh[off - 1] if off > 0 else 0.0, xh[off + cnt] if off + cnt < n else 0.0], np.float64)
        f = torch.zeros(1, device="cuda", dtype=torch.float64)
        assert lib.stochqn_b200_rosenbrock_grad(x.data_ptr(), g.data_ptr(), cnt, off, n, halo.data_ptr(), None) == 0
        assert lib.stochqn_b200_rosenbrock_fun(x.data_ptr(), cnt, off, n, halo.data_ptr(), f.data_ptr(), None) == 0
        assert np.max(np.abs(g.cpu().numpy() - gref[off:off + cnt])) <= 1e-12 * np.max(np.abs(gref))
        fsum += f.item()
    assert abs(fsum - Rosenbrock(n).fun(xh)) <= 1e-12 * abs(fsum)


def _logistic_np(X, y, sw, w, v, lam):
    """R/logistic.R:1-37 in NumPy."""
    p = 1.0 / (1.0 + np.exp(-(X @ w)))
    if sw is None:
        sw = np.ones(len(y))
    grad = X.T @ ((p - y) * sw) / sw.sum() + 2 * lam * w
    hv = X.T @ (p * (1 - p) * sw * (X @ v)) / sw.sum() + 2 * lam * v
    loss = np.sum(-(y * np.log(p) + (1 - y) * np.log(1 - p)) * sw) / sw.sum() + lam * (w @ w)
    return grad, hv, loss


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 2e-5)])
@pytest.mark.parametrize("shape", [(1, 1), (7, 3), (1000, 1001), (257, 64), (2000, 4097)])
@pytest.mark.parametrize("weighted", [False, True])
def test_logistic_grad_hessvec_loss(dtype, tol, shape, weighted):
    import torch
    abi = _lib.load(dtype)
    lib = abi.lib
    B, d = shape
    rng = np.random.default_rng(B * 7 + d)
    X = (rng.standard_normal((B, d)) / np.sqrt(d)).astype(dtype)
    w = rng.standard_normal(d).astype(dtype)
    v = rng.standard_normal(d).astype(dtype)
    y = (rng.random(B) < 0.5).astype(dtype)
    sw = (rng.random(B) + 0.5).astype(dtype) if weighted else None
    lam = 1e-5
    Xd, wd, vd, yd = _t(X, dtype), _t(w, dtype), _t(v, dtype), _t(y, dtype)
    swd = _t(sw, dtype) if weighted else None
    work = torch.empty(lib.stochqn_b200_logistic_work_size(B, d), device="cuda", dtype=torch.uint8)
    g = torch.empty_like(wd)
    hv = torch.empty_like(wd)
    loss = torch.zeros(1, device="cuda", dtype=torch.float64)
    swp = swd.data_ptr() if weighted else None
    assert lib.stochqn_b200_logistic_grad(Xd.data_ptr(), d, yd.data_ptr(), swp, B, d, wd.data_ptr(), lam, g.data_ptr(), work.data_ptr(), None) == 0
    assert lib.stochqn_b200_logistic_hess_vec(Xd.data_ptr(), d, yd.data_ptr(), swp, B, d, wd.data_ptr(), vd.data_ptr(), lam, hv.data_ptr(), work.data_ptr(), None) == 0
    assert lib.stochqn_b200_logistic_loss(Xd.data_ptr(), d, yd.data_ptr(), swp, B, d, wd.data_ptr(), lam, loss.data_ptr(), work.data_ptr(), None) == 0
    gr, hr, lr = _logistic_np(X.astype(np.float64), y.astype(np.float64), None if sw is None else sw.astype(np.float64),
                              w.astype(np.float64), v.astype(np.float64), lam)
    assert np.max(np.abs(g.cpu().numpy() - gr)) <= tol * max(np.max(np.abs(gr)), 1e-3) * 10
    assert np.max(np.abs(hv.cpu().numpy() - hr)) <= tol * max(np.max(np.abs(hr)), 1e-3) * 10
    assert abs(loss.item() - lr) <= max(tol, 1e-12) * abs(lr) * 10


def test_logistic_batch_is_a_row_range_view():
    """A batch is a row range of the resident matrix (ldx = full row length): no copy, same numbers."""
    import torch
    lib = _lib.load(np.float64).lib
    rng = np.random.default_rng(0)
    B, d = 500, 33
    X = rng.standard_normal((B, d))
    y = (rng.random(B) < 0.5).astype(np.float64)
    w = rng.standard_normal(d)
    Xd, yd, wd = _t(X, np.float64), _t(y, np.float64), _t(w, np.float64)
    g = torch.empty_like(wd)
    work = torch.empty(lib.stochqn_b200_logistic_work_size(B, d), device="cuda", dtype=torch.uint8)
    r0, r1 = 100, 300
    assert lib.stochqn_b200_logistic_grad(Xd.data_ptr() + r0 * d * 8, d, yd.data_ptr() + r0 * 8, None, r1 - r0, d, wd.data_ptr(), 1e-5,
                                          g.data_ptr(), work.data_ptr(), None) == 0
    gr, _, _ = _logistic_np(X[r0:r1], y[r0:r1], None, w, w, 1e-5)
    assert np.max(np.abs(g.cpu().numpy() - gr)) <= 1e-12 * np.max(np.abs(gr))


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 2e-5)])
@pytest.mark.parametrize("shape", [(1, 1), (7, 3), (1000, 1000), (257, 64), (600, 4096), (300, 5300)])
@pytest.mark.parametrize("weighted,fit_intercept", [(False, True), (True, False), (True, True)])
def test_logistic_sklearn_conventions(dtype, tol, shape, weighted, fit_intercept):
    """stochqn_b200_logistic_sk_* against the restatement of scikit-learn <= 1.0's _logistic_loss_and_grad /
    _logistic_grad_hess (oracle/logistic_sk_np.py): y in {-1,+1}, sums, intercept last and unpenalised.
    Covers the fused one-sweep kernel (ncols <= 5120) and the two-sweep form (5300 columns)."""
    import torch
    from stochqn_b200 import logistic as L
    from oracle import logistic_sk_np as LS

    B, d = shape
    rng = np.random.default_rng(B * 11 + d)
    X = (rng.standard_normal((B, d)) / np.sqrt(d)).astype(dtype)
    w = rng.standard_normal(d + fit_intercept).astype(dtype)
    v = rng.standard_normal(d + fit_intercept).astype(dtype)
    y = np.where(rng.random(B) < 0.5, 1.0, -1.0).astype(dtype)
    sw = (rng.random(B) + 0.5).astype(dtype) if weighted else None
    if sw is not None:
        sw /= sw.sum()
    alpha = 0.3
    Xd, wd, vd, yd = _t(X, dtype), _t(w, dtype), _t(v, dtype), _t(y, dtype)
    swd = _t(sw, dtype) if weighted else None
    g = L.grad_fun_bin(wd, Xd, yd, sample_weight=swd, reg_param=alpha).cpu().numpy()
    hv = L.hessvec_fun_bin(wd, vd, Xd, yd, sample_weight=swd, reg_param=alpha).cpu().numpy()
    loss = L.obj_fun_bin(wd, Xd, yd, sample_weight=swd, reg_param=alpha)
    f64 = lambda a: None if a is None else a.astype(np.float64)
    lr, gr = LS.logistic_loss_and_grad(f64(w), f64(X), f64(y), alpha, f64(sw))
    hr = LS.logistic_hess_vec(f64(w), f64(v), f64(X), f64(y), alpha, f64(sw))
    assert np.max(np.abs(g - gr)) <= tol * max(np.max(np.abs(gr)), 1e-3) * 10
    assert np.max(np.abs(hv - hr)) <= tol * max(np.max(np.abs(hr)), 1e-3) * 10
    assert abs(loss - lr) <= max(tol, 1e-12) * abs(lr) * 10
