"""Logistic regression fitted by the stochastic quasi-Newton optimizers, with every callback on the GPU.

Mirror of the reference's ``StochasticLogisticRegression`` (stochqn/_logistic.py:43-247: constructor arguments,
``fit`` / ``partial_fit`` / ``predict`` / ``predict_proba`` / ``coef_`` / ``intercept_``) on top of
``stochqn_b200.guided``.  The reference serves the optimizer's requests with scikit-learn's private
``_logistic_loss_and_grad`` / ``_logistic_grad_hess`` / ``_multinomial_loss_grad`` / ``_multinomial_grad_hess``
(stochqn/_logistic.py:3-30) on NumPy arrays; here the same closed forms are CUDA kernels of this library
(``stochqn_b200_logistic_sk_*`` - one fused sweep of the batch - and ``stochqn_b200_multinomial_*`` - two GEMMs,
tcgen05 tensor cores in the float build, with the soft-max / R-operator row kernel between them), the model
matrix is uploaded once and stays in HBM, and mini-batches are row-range views of it.

The module-level functions are the callbacks themselves, with the signatures the guided classes expect
(``f(w, X, y, sample_weight=None, reg_param=0)``), usable on their own with torch CUDA tensors:

    two classes :  grad_fun_bin, hessvec_fun_bin, obj_fun_bin, pred_fun_bin        y in {-1,+1}
    K classes   :  grad_fun_multi, hessvec_fun_mult, obj_fun_mult, pred_fun_mult    Y one-hot (n x K)
    R flavour   :  logistic_grad, logistic_hess_vec, logistic_loss                  y in {0,1}, means, R/logistic.R:1-37

There is no CPU implementation: without the CUDA library these raise.
"""
from __future__ import annotations

import ctypes as C
import warnings

import numpy as np

from . import _lib
from .guided import SQN, _is_sparse, _is_torch, _step_size_const, adaQN, oLBFGS


# ---- plumbing -------------------------------------------------------------------------------------------
_SCRATCH = {}


def _torch():
    import torch
    return torch


def _scratch(device, nbytes):
    """Per-device scratch buffer for the kernels' partial sums (grown on demand, reused between calls)."""
    torch = _torch()
    key = (device.type, device.index)
    buf = _SCRATCH.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device=device)
        _SCRATCH[key] = buf
    return buf


def _abi_for(t):
    torch = _torch()
    if t.dtype == torch.float64:
        return _lib.load(np.float64)
    if t.dtype == torch.float32:
        return _lib.load(np.float32)
    raise ValueError("the logistic callbacks take float64 or float32 tensors, got %s" % (t.dtype,))


def _rows(X):
    """(tensor, leading dimension) of a row-major device matrix; copies only when the rows are not contiguous."""
    if not (_is_torch(X) and X.is_cuda):
        raise TypeError("the device callbacks need torch CUDA tensors (dense); got %s" % type(X).__name__)
    if X.dim() != 2:
        raise ValueError("X must be 2-dimensional")
    if X.stride(1) != 1 or (X.shape[0] > 1 and X.stride(0) < X.shape[1]):
        X = X.contiguous()
    return X, (X.stride(0) if X.shape[0] > 1 else X.shape[1])


def _vec(a, like, name):
    if a is None:
        return None
    torch = _torch()
    if not _is_torch(a):
        a = torch.as_tensor(np.asarray(a), device=like.device)
    a = a.to(device=like.device, dtype=like.dtype).reshape(-1)
    return a if a.is_contiguous() else a.contiguous()


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return C.c_void_p(_torch().cuda.current_stream().cuda_stream)


def _check(rc, abi, what):
    if rc != 0:
        raise RuntimeError("%s failed: %s" % (what, _lib.last_error(abi)))


# ---- two classes, scikit-learn conventions (stochqn/_logistic.py:23-36) --------------------------------------
def _bin_call(kind, w, v, X, y, sample_weight, reg_param):
    torch = _torch()
    X, ldx = _rows(X)
    abi = _abi_for(X)
    lib = abi.lib
    nrows, ncols = X.shape
    w = _vec(w, X, "w")
    icpt = int(w.numel() == ncols + 1)
    if w.numel() not in (ncols, ncols + 1):
        raise ValueError("w must have n_features or n_features + 1 entries")
    y = _vec(y, X, "y")
    sw = _vec(sample_weight, X, "sample_weight")
    work = _scratch(X.device, lib.stochqn_b200_logistic_work_size(nrows, ncols))
    alpha = float(reg_param)
    if kind == "grad":
        out = torch.empty_like(w)
        _check(lib.stochqn_b200_logistic_sk_grad(X.data_ptr(), ldx, y.data_ptr(), _ptr(sw), nrows, ncols, icpt,
                                                 w.data_ptr(), alpha, out.data_ptr(), work.data_ptr(), _stream()), abi, "logistic gradient")
        return out
    if kind == "hvp":
        v = _vec(v, X, "v")
        out = torch.empty_like(w)
        _check(lib.stochqn_b200_logistic_sk_hess_vec(X.data_ptr(), ldx, y.data_ptr(), _ptr(sw), nrows, ncols, icpt,
                                                     w.data_ptr(), v.data_ptr(), alpha, out.data_ptr(), work.data_ptr(),
                                                     _stream()), abi, "logistic Hessian-vector product")
        return out
    loss = torch.empty(1, dtype=torch.float64, device=X.device)
    _check(lib.stochqn_b200_logistic_sk_loss(X.data_ptr(), ldx, y.data_ptr(), _ptr(sw), nrows, ncols, icpt,
                                             w.data_ptr(), alpha, loss.data_ptr(), work.data_ptr(), _stream()), abi, "logistic loss")
    return float(loss.item())


def grad_fun_bin(w, X, y, sample_weight=None, reg_param=0):
    return _bin_call("grad", w, None, X, y, sample_weight, reg_param)


def hessvec_fun_bin(w, v, X, y, sample_weight=None, reg_param=0):
    return _bin_call("hvp", w, v, X, y, sample_weight, reg_param)


def obj_fun_bin(w, X, y, sample_weight=None, reg_param=0):
    return _bin_call("loss", w, None, X, y, sample_weight, reg_param)


def pred_fun_bin(w, X):
    """sigmoid(X w[:d] (+ intercept)) - stochqn/_logistic.py:31-36.  A GEMV: plain torch."""
    torch = _torch()
    d = X.shape[1]
    z = X @ w[:d]
    if w.shape[0] != d:
        z = z + w[-1]
    return torch.sigmoid(z).reshape(-1)


# ---- K classes, scikit-learn conventions (stochqn/_logistic.py:7-21) ------------------------------------------
def _mult_call(kind, w, v, X, Y, sample_weight, reg_param):
    torch = _torch()
    X, ldx = _rows(X)
    abi = _abi_for(X)
    lib = abi.lib
    nrows, nfeat = X.shape
    if not _is_torch(Y):
        Y = torch.as_tensor(np.asarray(Y), device=X.device)
    Y = Y.to(device=X.device, dtype=X.dtype)
    Y, ldy = _rows(Y)
    K = Y.shape[1]
    w = _vec(w, X, "w")
    if w.numel() not in (K * nfeat, K * (nfeat + 1)):
        raise ValueError("w must have n_classes * (n_features [+ 1]) entries")
    icpt = int(w.numel() == K * (nfeat + 1))
    sw = _vec(sample_weight, X, "sample_weight")
    work = _scratch(X.device, lib.stochqn_b200_multinomial_work_size(nrows, nfeat, K))
    alpha = float(reg_param)
    if kind == "hvp":
        v = _vec(v, X, "v")
        out = torch.empty_like(w)
        _check(lib.stochqn_b200_multinomial_hess_vec(X.data_ptr(), ldx, Y.data_ptr(), ldy, None, _ptr(sw), nrows, nfeat, K,
                                                     icpt, w.data_ptr(), v.data_ptr(), alpha, out.data_ptr(), work.data_ptr(),
                                                     _stream()), abi, "multinomial Hessian-vector product")
        return out
    grad = torch.empty_like(w) if kind == "grad" else None
    loss = torch.empty(1, dtype=torch.float64, device=X.device) if kind == "loss" else None
    _check(lib.stochqn_b200_multinomial_loss_grad(X.data_ptr(), ldx, Y.data_ptr(), ldy, None, _ptr(sw), nrows, nfeat, K, icpt,
                                                  w.data_ptr(), alpha, _ptr(grad), _ptr(loss), work.data_ptr(), _stream()),
           abi, "multinomial loss / gradient")
    return grad if kind == "grad" else float(loss.item())


def grad_fun_multi(w, X, y, sample_weight=None, reg_param=0):
    return _mult_call("grad", w, None, X, y, sample_weight, reg_param)


def obj_fun_mult(w, X, y, sample_weight=None, reg_param=0):
    return _mult_call("loss", w, None, X, y, sample_weight, reg_param)


def hessvec_fun_mult(w, v, X, y, sample_weight=None, reg_param=0):
    return _mult_call("hvp", w, v, X, y, sample_weight, reg_param)


def pred_fun_mult(w, X, nclasses):
    """Per-class sigmoid(X W' (+ b)) as an (n_samples, n_classes) matrix - stochqn/_logistic.py:14-21, which applies the
    logistic function (not the soft-max) to the scores; the arg-max, i.e. ``predict``, is the same either way.
    (Without an intercept the reference returns the transposed matrix, ``w.dot(X.T)``; here the orientation is
    (n_samples, n_classes) in both cases.)"""
    torch = _torch()
    W = w.reshape(nclasses, -1)
    d = X.shape[1]
    z = X @ W[:, :d].T
    if W.shape[1] != d:
        z = z + W[:, -1].reshape(1, -1)
    return torch.sigmoid(z)


# ---- R conventions (R/logistic.R:1-37): y in {0,1}, means, intercept as a column of X ------------------------------
def _r_call(kind, w, v, X, y, sample_weight, reg_param):
    torch = _torch()
    X, ldx = _rows(X)
    abi = _abi_for(X)
    lib = abi.lib
    nrows, ncols = X.shape
    w, y, sw = _vec(w, X, "w"), _vec(y, X, "y"), _vec(sample_weight, X, "sample_weight")
    work = _scratch(X.device, lib.stochqn_b200_logistic_work_size(nrows, ncols))
    lam = float(reg_param)
    if kind == "grad":
        out = torch.empty_like(w)
        _check(lib.stochqn_b200_logistic_grad(X.data_ptr(), ldx, y.data_ptr(), _ptr(sw), nrows, ncols, w.data_ptr(), lam,
                                              out.data_ptr(), work.data_ptr(), _stream()), abi, "logistic gradient")
        return out
    if kind == "hvp":
        v = _vec(v, X, "v")
        out = torch.empty_like(w)
        _check(lib.stochqn_b200_logistic_hess_vec(X.data_ptr(), ldx, y.data_ptr(), _ptr(sw), nrows, ncols, w.data_ptr(),
                                                  v.data_ptr(), lam, out.data_ptr(), work.data_ptr(), _stream()), abi,
               "logistic Hessian-vector product")
        return out
    loss = torch.empty(1, dtype=torch.float64, device=X.device)
    _check(lib.stochqn_b200_logistic_loss(X.data_ptr(), ldx, y.data_ptr(), _ptr(sw), nrows, ncols, w.data_ptr(), lam,
                                          loss.data_ptr(), work.data_ptr(), _stream()), abi, "logistic loss")
    return float(loss.item())


def logistic_grad(w, X, y, sample_weight=None, reg_param=1e-5):
    return _r_call("grad", w, None, X, y, sample_weight, reg_param)


def logistic_hess_vec(w, v, X, y, sample_weight=None, reg_param=1e-5):
    return _r_call("hvp", w, v, X, y, sample_weight, reg_param)


def logistic_loss(w, X, y, sample_weight=None, reg_param=1e-5):
    return _r_call("loss", w, None, X, y, sample_weight, reg_param)


# ---- the estimator ----------------------------------------------------------------------------------------------
class StochasticLogisticRegression:
    """Logistic regression fit with a stochastic quasi-Newton optimizer (reference: stochqn/_logistic.py:43-247).

    Parameters (same names and defaults as the reference)
    ----------
    reg_param : float          strength of the l2 penalty (the loss is an average over observations)
    fit_intercept : bool       add an intercept (stored last, not penalised)
    random_state : int         seed of the starting point (``np.random.normal``) and of shuffling / splitting
    optimizer : str            'oLBFGS', 'SQN' or 'adaQN'
    step_size : float          initial step size
    valset_frac : float/None   share of the data held out to monitor the objective after each epoch
    verbose : bool
    optimizer_kwargs           passed on to the optimizer (``stochqn_b200.guided``), e.g. ``use_float=True``
    device                     (new) CUDA device of the model; default: the device of ``X`` if it is a CUDA tensor,
                               else the current CUDA device

    ``X`` may be a NumPy array (uploaded once) or a torch CUDA tensor; labels ``y`` are {-1,+1} for two classes
    (the scikit-learn functions the reference calls assume that coding) or a one-hot matrix for several.
    """

    def __init__(self, reg_param=1e-3, fit_intercept=True, random_state=1, optimizer="SQN", step_size=1e-1,
                 valset_frac=0.1, verbose=False, device=None, **optimizer_kwargs):
        assert optimizer in ["oLBFGS", "SQN", "adaQN"]
        assert isinstance(step_size, float) and step_size > 0
        assert isinstance(reg_param, float) and reg_param >= 0
        optimizer_kwargs["step_size"] = step_size
        optimizer_kwargs["valset_frac"] = valset_frac
        optimizer_kwargs["verbose"] = verbose
        self.optimizer_name = optimizer
        self.optimizer = None
        self.optimizer_kwargs = optimizer_kwargs
        self.reg_param = reg_param
        self.nclasses = None
        self._is_mult = None
        self.fit_intercept = bool(fit_intercept)
        self.is_fitted = False
        self.random_state = random_state
        self.device = device
        self._numpy_io = True

    # ---- fitted attributes ----------------------------------------------------------------------------------
    def _out(self, t):
        return t.detach().cpu().numpy() if self._numpy_io else t

    @property
    def coef_(self):
        if not self.is_fitted:
            return None
        x = self.optimizer.x
        if self._is_mult:
            W = x.reshape(self.nclasses, -1)
            return self._out(W[:, :-1] if self.fit_intercept else W)
        return self._out(x[:-1] if self.fit_intercept else x)

    @property
    def intercept_(self):
        if not self.is_fitted:
            return None
        x = self.optimizer.x
        if self._is_mult:
            if self.fit_intercept:
                return self._out(x.reshape(self.nclasses, -1)[:, -1])
            return np.zeros(self.nclasses)
        return float(x[-1].item()) if self.fit_intercept else 0.0

    # ---- predictions ----------------------------------------------------------------------------------------
    def _to_dev(self, a, dtype=None):
        torch = _torch()
        if _is_sparse(a):
            raise TypeError("sparse inputs are not supported by the device callbacks; pass a dense array")
        x = self.optimizer.x if self.optimizer is not None else None
        dev = x.device if x is not None else torch.device(self.device if self.device is not None else "cuda")
        dt = dtype if dtype is not None else (x.dtype if x is not None else None)
        if not _is_torch(a):
            a = torch.as_tensor(np.ascontiguousarray(a))
        return a.to(device=dev, dtype=dt)

    def predict_proba(self, X):
        """Class probabilities, (n_samples, n_classes) (two classes: columns [1 - p, p])."""
        torch = _torch()
        as_numpy = not _is_torch(X)
        Xd = self._to_dev(X)
        if self._is_mult:
            out = pred_fun_mult(self.optimizer.x, Xd, self.nclasses)
        else:
            p = pred_fun_bin(self.optimizer.x, Xd).reshape(-1, 1)
            out = torch.cat([1 - p, p], dim=1)
        return out.cpu().numpy() if as_numpy else out

    def predict(self, X):
        """Predicted class: index of the largest score (several classes) or 0 / 1 (two classes)."""
        torch = _torch()
        as_numpy = not _is_torch(X)
        Xd = self._to_dev(X)
        if self._is_mult:
            out = torch.argmax(pred_fun_mult(self.optimizer.x, Xd, self.nclasses), dim=1)
        else:
            out = (pred_fun_bin(self.optimizer.x, Xd) >= .5).to(torch.uint8)
        return out.cpu().numpy() if as_numpy else out

    # ---- fitting ----------------------------------------------------------------------------------------------
    def _check_fit_inp(self, X, y, sample_weight):
        torch = _torch()
        assert X.shape[0] == y.shape[0]
        use_float = bool(self.optimizer_kwargs.get("use_float", False))
        dt = torch.float32 if use_float else torch.float64
        self._numpy_io = not _is_torch(X)
        if self.optimizer is None and self.device is None and _is_torch(X) and X.is_cuda:
            self.device = X.device
        if _is_sparse(y):
            warnings.warn("'StochasticLogisticRegression' only supports dense arrays for 'y', will cast the array.")
            y = np.array(y.todense())
        Xd = self._to_dev(X, dt)
        yd = self._to_dev(y, dt)
        if sample_weight is None:
            sw = torch.ones(X.shape[0], dtype=dt, device=Xd.device)
        else:
            sw = self._to_dev(sample_weight, dt).reshape(-1).clone()
        assert sw.shape[0] == X.shape[0]
        sw = sw / sw.sum()          # the callbacks compute sums, not means (stochqn/_logistic.py:167)
        return Xd, yd, sw

    def _initialize_optimizer(self, X, y):
        if self.optimizer is not None:
            return
        torch = _torch()
        if y.dim() == 1:
            self._is_mult, self.nclasses = False, 2
            funs = dict(obj_fun=obj_fun_bin, grad_fun=grad_fun_bin, pred_fun=pred_fun_bin)
            hv = hessvec_fun_bin
        else:
            self._is_mult, self.nclasses = True, y.shape[1]
            funs = dict(obj_fun=obj_fun_mult, grad_fun=grad_fun_multi, pred_fun=pred_fun_mult)
            hv = hessvec_fun_mult
        np.random.seed(self.random_state)
        w0 = np.random.normal(size=(X.shape[1] + self.fit_intercept) * (y.shape[1] if self._is_mult else 1))
        w0 = torch.as_tensor(w0, device=X.device).to(X.dtype)
        if self.optimizer_name == "oLBFGS":
            self.optimizer = oLBFGS(x0=w0, **funs, **self.optimizer_kwargs)
        elif self.optimizer_name == "SQN":
            self.optimizer = SQN(x0=w0, hess_vec_fun=hv, **funs, **self.optimizer_kwargs)
        else:
            self.optimizer = adaQN(x0=w0, **funs, **self.optimizer_kwargs)

    def fit(self, X, y, sample_weight=None):
        """Fit the model in stochastic batches (``batches_per_epoch`` / ``nepochs`` of the optimizer)."""
        X, y, sample_weight = self._check_fit_inp(X, y, sample_weight)
        self._initialize_optimizer(X, y)
        self.optimizer.fit(X, y, sample_weight, {"reg_param": self.reg_param})
        self.is_fitted = True
        return self

    def partial_fit(self, X, y, sample_weight=None, classes=None, decr_step_size=False):
        """Update the model with one batch; `classes` is ignored (scikit-learn API compatibility).  The step size
        follows the optimizer's ``decr_step_size`` schedule only when `decr_step_size` is true."""
        X, y, sample_weight = self._check_fit_inp(X, y, sample_weight)
        self._initialize_optimizer(X, y)
        if decr_step_size:
            self.optimizer.partial_fit(X, y, sample_weight, {"reg_param": self.reg_param})
        else:
            keep = self.optimizer.decr_step_size
            self.optimizer.decr_step_size = _step_size_const
            try:
                self.optimizer.partial_fit(X, y, sample_weight, {"reg_param": self.reg_param})
            finally:
                self.optimizer.decr_step_size = keep
        self.is_fitted = True
        return self
