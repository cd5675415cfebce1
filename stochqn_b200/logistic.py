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


def _csr_to_dense_on_device(a, dev, dt):
    """A sparse model matrix (scipy sparse, any format, or a torch sparse-CSR tensor - stochqn/_logistic.py:155 keeps CSR inputs)
    expanded ON THE DEVICE into the dense row-major matrix the bundled kernels stream: the three CSR arrays are uploaded and
    one kernel (stochqn_b200_csr_to_dense) zero-fills and scatters every row.  The dense copy must fit in device memory."""
    torch = _torch()
    if _is_torch(a):
        indptr, indices, data = a.crow_indices(), a.col_indices(), a.values()
        nrows, ncols = a.shape
    else:
        from scipy.sparse import csr_matrix
        a = csr_matrix(a)
        a.sum_duplicates()
        indptr, indices, data = torch.as_tensor(a.indptr.astype(np.int64)), torch.as_tensor(a.indices.astype(np.int64)), torch.as_tensor(a.data)
        nrows, ncols = a.shape
    need = int(nrows) * int(ncols) * (8 if dt == torch.float64 else 4)
    free, _total = torch.cuda.mem_get_info(dev)
    if need > 0.9 * free:
        raise MemoryError("the dense copy of the sparse model matrix (%d x %d, %.1f GB) does not fit in device memory (%.1f GB free); "
                          "the bundled device callbacks are dense kernels" % (nrows, ncols, need / 1e9, free / 1e9))
    indptr = indptr.to(device=dev, dtype=torch.int64).contiguous()
    indices = indices.to(device=dev, dtype=torch.int64).contiguous()
    data = data.to(device=dev, dtype=dt).contiguous()
    out = torch.empty((int(nrows), int(ncols)), device=dev, dtype=dt)
    bad = torch.zeros(1, device=dev, dtype=torch.int32)
    abi = _lib.load(np.float64 if dt == torch.float64 else np.float32)
    with torch.cuda.device(dev):
        _check(abi.lib.stochqn_b200_csr_to_dense(indptr.data_ptr(), indices.data_ptr(), data.data_ptr() if data.numel() else None, 0, int(nrows),
                                                 int(ncols), out.data_ptr(), int(ncols), bad.data_ptr(), _stream()), abi, "csr_to_dense")
    if int(bad.item()):
        raise ValueError("the sparse matrix holds a column index outside [0, %d)" % ncols)
    return out


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


# ---- the request loop of a mini-batch inside the library ----------------------------------------------------------
class _NativeLoop:
    """Mixin for the guided classes: when the data are device-resident row ranges and the callbacks are the bundled
    ones, ``_fit_batch`` is ONE call into the library (``stochqn_b200_fit_batch``, include/stochqn_b200.h) instead of
    one Python round trip per request.  Anything the native loop cannot serve exactly as the reference would (a long
    batch that is not a row range, a step size that changes inside the loop, verbose event messages) goes through
    the generic Python loop of ``stochqn_b200.guided`` - the results are the same either way
    (tests/test_gpu_logistic_estimator.py runs both)."""

    _native_model = None          # (model id, fit_intercept, nclasses) set by the estimator
    native_batches = 0            # mini-batches served natively (diagnostics / tests)
    use_native_loop = True

    def _rows_struct(self, X, y, w, keep):
        if X is None or not (_is_torch(X) and X.is_cuda and X.dim() == 2 and X.stride(1) == 1 and X.dtype == self.x.dtype):
            return None
        if not (_is_torch(y) and y.is_cuda and y.dtype == X.dtype and y.shape[0] == X.shape[0]):
            return None
        if y.dim() == 2:
            if y.stride(1) != 1:
                return None
            ldy = y.stride(0) if y.shape[0] > 1 else y.shape[1]
        else:
            if y.shape[0] > 1 and y.stride(0) != 1:
                return None
            ldy = 1
        if w is not None:
            if not (_is_torch(w) and w.is_cuda and w.dtype == X.dtype and w.dim() == 1 and (w.shape[0] <= 1 or w.stride(0) == 1)):
                return None
        keep.extend((X, y, w))
        ldx = X.stride(0) if X.shape[0] > 1 else X.shape[1]
        return _lib.Rows(X.data_ptr(), ldx, y.data_ptr(), ldy, w.data_ptr() if w is not None else None, X.shape[0])

    def _peek_stash(self):
        """The stored batches as ONE row range (X, y, w) without emptying the stash, or None when they are not
        consecutive rows of the same arrays (or mix weighted and unweighted batches)."""
        from .guided import _merge_adjacent
        st = self._stash
        if not len(st):
            return None
        weighted = [w is not None for w in st.w]
        if any(weighted) and not all(weighted):
            return None
        one = len(st) == 1
        X = st.X[0] if one else _merge_adjacent(st.X)
        y = st.y[0] if one else _merge_adjacent(st.y)
        w = None if not weighted[0] else (st.w[0] if one else _merge_adjacent(st.w))
        if X is None or y is None or (weighted[0] and w is None):
            return None
        return X, y, w

    def _fit_batch(self, X_batch, y_batch, w_batch, additional_kwargs, is_user_batch=False,
                   X_full=None, y_full=None, w_full=None, X_val=None, y_val=None, w_val=None, batch=None):
        generic = super()._fit_batch
        args = (X_batch, y_batch, w_batch, additional_kwargs)
        kw = dict(is_user_batch=is_user_batch, X_full=X_full, y_full=y_full, w_full=w_full, X_val=X_val, y_val=y_val,
                  w_val=w_val, batch=batch)
        free = self.optimizer
        const_step = (not is_user_batch) or self.decr_step_size is _step_size_const
        if (not self.use_native_loop or self._native_model is None or self.verbose or not const_step
                or not getattr(free, "initialized", False) or not hasattr(free, "_ws") or not _is_torch(self.x)
                or set(additional_kwargs) - {"reg_param"}):
            return generic(*args, **kw)
        keep = []
        rb = self._rows_struct(X_batch, y_batch, w_batch, keep)
        if rb is None:
            return generic(*args, **kw)
        # the long batch as ONE row range, when the reference's rule yields one without touching the stash
        rl, from_stash = None, False
        if self.optimizer_name != "oLBFGS":
            if is_user_batch:
                view = self._peek_stash()
                if view is not None:
                    rl = self._rows_struct(view[0], view[1], view[2], keep)
                    from_stash = rl is not None
            elif X_full is not None:
                first, last, diff = self._fit_long_rows(X_full.shape[0], batch)
                if diff == 0:
                    rl = self._rows_struct(X_full[first:last], y_full[first:last],
                                           w_full[first:last] if w_full is not None else None, keep)
        rv = self._rows_struct(X_val, y_val, w_val, keep) if X_val is not None else None
        if X_val is not None and rv is None:
            return generic(*args, **kw)

        abi = free._abi
        model_id, icpt, nclasses = self._native_model
        ncols = X_batch.shape[1]
        nmax = max(r.nrows for r in (rb, rl, rv) if r is not None)
        if model_id == 2:
            work = _scratch(X_batch.device, abi.lib.stochqn_b200_multinomial_work_size(nmax, ncols, nclasses))
        else:
            work = _scratch(X_batch.device, abi.lib.stochqn_b200_logistic_work_size(nmax, ncols))
        M = abi.Model(model_id, icpt, ncols, nclasses, float(additional_kwargs.get("reg_param", 0)), work.data_ptr())
        task = C.c_int({v: k for k, v in _TASK_CODES.items()}[self.req["task"]])
        if not hasattr(free, "_req_vec"):
            free._req_vec = C.c_void_p()
        rep = _lib.FitReport()
        step = self.decr_step_size(self.step_size, self.niter if is_user_batch else self.epoch)
        rc = abi.lib.stochqn_b200_fit_batch(free._ws, self.x.data_ptr(), float(step), C.byref(M), C.byref(rb),
                                            C.byref(rl) if rl is not None else None, C.byref(rv) if rv is not None else None,
                                            C.byref(task), C.byref(free._req), C.byref(free._req_vec), C.byref(rep))
        if rc < 0:
            raise RuntimeError("stochqn_b200_fit_batch failed: " + _lib.last_error(abi))
        if rep.long_batch_used and from_stash:
            self._stash.clear()               # the reference empties its stash when it serves a long batch
        n = self.x.shape[0]
        at = free._wrap_req(free._req.value, self.x, n)
        if task.value == 104:
            at = (at, free._wrap_req(free._req_vec.value, self.x, n))
        from .optimizers import info_dct
        self.req = {"task": _TASK_CODES[task.value], "requested_on": at,
                    "info": {"x_changed_in_run": bool(rep.x_changed), "iteration_number": self.niter,
                             "iteration_info": info_dct[rep.last_info]}}
        if rc == 1:                            # a request the native loop could not serve: the generic loop carries on
            return generic(*args, **kw)
        self.native_batches = self.native_batches + 1
        if self.callback_iter is not None:
            self.callback_iter(self.x, **self.kwargs_cb)


_TASK_CODES = {101: "calc_grad", 102: "calc_grad_same_batch", 103: "calc_grad_big_batch", 104: "calc_hess_vec",
               105: "calc_fun_val_batch"}


class _oLBFGS_native(_NativeLoop, oLBFGS):
    pass


class _SQN_native(_NativeLoop, SQN):
    pass


class _adaQN_native(_NativeLoop, adaQN):
    pass


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
        self.native_loop = True       # False: serve every request from Python (same results; for comparison / debugging)
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
        x = self.optimizer.x if self.optimizer is not None else None
        dev = x.device if x is not None else torch.device(self.device if self.device is not None else "cuda")
        dt = dtype if dtype is not None else (x.dtype if x is not None else None)
        if _is_sparse(a) or (_is_torch(a) and a.layout == torch.sparse_csr):
            return _csr_to_dense_on_device(a, dev, dt if dt is not None else torch.float64)
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
            self.optimizer = _oLBFGS_native(x0=w0, **funs, **self.optimizer_kwargs)
        elif self.optimizer_name == "SQN":
            self.optimizer = _SQN_native(x0=w0, hess_vec_fun=hv, **funs, **self.optimizer_kwargs)
        else:
            self.optimizer = _adaQN_native(x0=w0, **funs, **self.optimizer_kwargs)
        # the request loop of a mini-batch runs inside the library (stochqn_b200_fit_batch) whenever it can
        self.optimizer._native_model = (2 if self._is_mult else 1, int(self.fit_intercept), self.nclasses if self._is_mult else 0)
        self.optimizer.use_native_loop = self.native_loop

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
