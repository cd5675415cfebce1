"""Guided-mode optimizers over the CUDA free-mode classes.

Host-side mirror of the reference's scikit-learn-style layer (stochqn/_optimizers.py:31-413 ``_StochQN`` with
``fit`` / ``partial_fit`` / ``predict`` / ``get_x``; constructors at 508-521 ``oLBFGS``, 626-646 ``SQN``,
762-783 ``adaQN``; R twin: R/optimizers_guided.R:26-111, R/helpers.R:146-191): same class names, constructor
arguments and defaults, the same batch schedule, long-batch rule, step-size schedules, shuffling seeds,
validation split and stopping rule - so a script written against the reference runs unchanged.  What is new:

* ``x0``, ``X``, ``y``, ``sample_weight`` may be **torch CUDA tensors**.  The optimizer state then lives in HBM
  (``stochqn_b200.optimizers``), ``requested_on`` handed to the user's callbacks is a zero-copy CUDA view,
  mini-batches are row-range views of the resident matrix, and the epoch shuffle is one device gather.
* The "long batch" (the union of the last ``bfgs_upd_freq`` mini-batches on which SQN / adaQN want a
  Hessian-vector product, a big-batch gradient or a function value) is handed over as **one row-range view**
  whenever the stored batches are adjacent rows of the same array - which is always the case inside ``fit`` and
  for users who stream consecutive slices into ``partial_fit`` - instead of the reference's ``np.r_`` / ``rbind``
  copy (stochqn/_optimizers.py:80-107); only genuinely scattered batches are concatenated.

NumPy inputs keep working exactly as in the reference (the library stages them through the GPU itself).
There is no CPU implementation behind these classes.
"""
from __future__ import annotations

import warnings

import numpy as np

from .optimizers import adaQN_free, oLBFGS_free, SQN_free


# ---- step-size schedules (stochqn/_optimizers.py:22-26) ---------------------------------------------
def _step_size_sqrt(initial_step_size, iteration_num):
    return initial_step_size / np.sqrt(iteration_num + 1)


def _step_size_const(initial_step_size, iteration_num):
    return initial_step_size


# ---- array helpers that work for NumPy arrays and torch tensors alike ----------------------------------
def _is_torch(a):
    return type(a).__module__.split(".")[0] == "torch"


def _is_sparse(a):
    return hasattr(a, "tocsr") and hasattr(a, "nnz")


def _take_rows(a, order):
    """a[order] for a host permutation `order`; torch tensors are gathered on their own device."""
    if a is None:
        return None
    if _is_torch(a):
        import torch
        return a[torch.as_tensor(order, device=a.device, dtype=torch.long)]
    return a[order]


def _row_view_span(a):
    """(storage id, first byte, bytes per row, rows) when `a` is a row-contiguous dense block, else None."""
    if a is None or _is_sparse(a):
        return None
    if _is_torch(a):
        if a.dim() == 0 or not a.is_contiguous():
            return None
        row_bytes = (int(np.prod(a.shape[1:], dtype=np.int64))) * a.element_size()
        return (("t", a.device.type, a.device.index, a.untyped_storage().data_ptr(), a.dtype, tuple(a.shape[1:])),
                a.data_ptr(), row_bytes, a.shape[0])
    if isinstance(a, np.ndarray):
        if a.ndim == 0 or not a.flags["C_CONTIGUOUS"]:
            return None
        base = a
        while isinstance(base.base, np.ndarray):
            base = base.base
        row_bytes = int(np.prod(a.shape[1:], dtype=np.int64)) * a.itemsize
        return (("n", base.ctypes.data, a.dtype.str, tuple(a.shape[1:])), a.ctypes.data, row_bytes, a.shape[0])
    return None


def _merge_adjacent(blocks):
    """One view over all `blocks` when they are consecutive row ranges of the same array (zero copy), else None."""
    spans = [_row_view_span(b) for b in blocks]
    if any(s is None for s in spans):
        return None
    key, start, row_bytes, _ = spans[0]
    if row_bytes == 0:
        return None
    nxt = start
    total = 0
    for k, p, rb, rows in spans:
        if k != key or rb != row_bytes or p != nxt:
            return None
        nxt = p + rows * rb
        total += rows
    first = blocks[0]
    shape = (total,) + tuple(first.shape[1:])
    if _is_torch(first):
        return first.as_strided(shape, first.stride())
    return np.lib.stride_tricks.as_strided(first, shape=shape, strides=first.strides, writeable=False)


def _stack_rows(blocks):
    """Rows of all `blocks` as one array: a view when they are adjacent, a concatenation otherwise
    (mixed sparse / dense lists are forced to dense, as stochqn/_optimizers.py:80-90 does)."""
    if len(blocks) == 1:
        return blocks[0]
    merged = _merge_adjacent(blocks)
    if merged is not None:
        return merged
    n_sparse = sum(1 for b in blocks if _is_sparse(b))
    if n_sparse == len(blocks):
        from scipy.sparse import vstack
        return vstack(blocks)
    if n_sparse:
        warnings.warn("When passing mixed batches of sparse and non-sparse data, these are forced to dense.")
        blocks = [np.asarray(b.todense()) if _is_sparse(b) else b for b in blocks]
    if any(_is_torch(b) for b in blocks):
        import torch
        dev = next(b.device for b in blocks if _is_torch(b))
        return torch.cat([b if _is_torch(b) else torch.as_tensor(b, device=dev) for b in blocks], dim=0)
    return np.concatenate(blocks, axis=0)


class _RowStash:
    """The limited-memory container of mini-batches that make up the next long batch
    (stochqn/_optimizers.py:92-112 ``_get_stored_batch`` / ``_reset_saved_batch``)."""

    def __init__(self):
        self.clear()

    def clear(self):
        self.X, self.y, self.w = [], [], []

    def push(self, X, y, w):
        self.X.append(X)
        self.y.append(y)
        self.w.append(w)

    def __len__(self):
        return len(self.X)

    def pop_all(self):
        if not self.X:
            raise ValueError("Unexpected error: a long batch was requested but no batches are stored.")
        X_long, y_long = _stack_rows(self.X), _stack_rows(self.y)
        missing = sum(1 for w in self.w if w is None)
        if missing == len(self.w):
            w_long = None
        else:
            if missing:
                warnings.warn("Passed batches with and without sample weights, missing weights will be set to 1.")
                ws = []
                for Xb, wb in zip(self.X, self.w):
                    if wb is None:
                        like = next(w for w in self.w if w is not None)
                        if _is_torch(like):
                            import torch
                            wb = torch.ones(Xb.shape[0], dtype=like.dtype, device=like.device)
                        else:
                            wb = np.ones(Xb.shape[0], dtype=getattr(like, "dtype", np.float64))
                    ws.append(wb)
                self.w = ws
            w_long = _stack_rows(self.w)
        self.clear()
        return X_long, y_long, w_long


class _StochQN:
    """Behaviour shared by the three guided optimizers (reference: stochqn/_optimizers.py:31-413)."""

    # ---- input checks -------------------------------------------------------------------------------
    def _check_fit_inputs(self, X, y, sample_weight, additional_kwargs=None, check_sp=False):
        assert X.shape[0] > 0
        assert X.shape[0] == y.shape[0]
        if sample_weight is not None:
            assert sample_weight.shape[0] == X.shape[0]
        if additional_kwargs is None:
            additional_kwargs = dict()
        assert isinstance(additional_kwargs, dict)
        if check_sp:
            X, y = self._to_csr(X), self._to_csr(y)
            sample_weight = self._to_csr(sample_weight) if sample_weight is not None else None
        return X, y, sample_weight, additional_kwargs

    @staticmethod
    def _to_csr(a):
        if _is_sparse(a) and a.format != "csr":
            warnings.warn("'.fit' method only supports sparse CSR matrices. Sparse inputs will be cast to CSR.")
            from scipy.sparse import csr_matrix
            a = csr_matrix(a)
        return a

    # ---- construction -------------------------------------------------------------------------------
    def _add_common_attributes(self, x0, batches_per_epoch, step_size, grad_fun, obj_fun, pred_fun, decr_step_size,
                               callback_epoch, callback_iter, valset_frac, tol, nepochs, kwargs_cb, random_state,
                               shuffle_data, verbose, use_grad_diff, use_float):
        assert isinstance(batches_per_epoch, int) and batches_per_epoch > 0
        assert step_size > 0
        if decr_step_size == "auto":
            decr_step_size = _step_size_sqrt
        elif decr_step_size is None:
            decr_step_size = _step_size_const
        elif not callable(decr_step_size):
            raise ValueError("'decr_step_size' must be a function taking as input the initial step size and the "
                             "iteration number, starting at zero.")
        for cb in (callback_epoch, callback_iter):
            if cb is not None and not callable(cb):
                raise ValueError("Callback must be a function taking as argument the values of 'x' and additional "
                                 "keyword arguments, or 'None'")
        if not callable(grad_fun):
            raise ValueError("'grad_fun' must be a function that takes as argument the variables values, X, y, "
                             "sample_weight, and additional keyword arguments.")
        if pred_fun is not None and not callable(pred_fun):
            raise ValueError("'pred_fun' must be a function that takes as argument the variables values and X, or 'None'.")
        if valset_frac is not None:
            assert 0 < valset_frac < 1
            assert tol > 0
            if not callable(obj_fun):
                raise ValueError("'obj_fun' must be a function that takes as argument the variables values, X, y, "
                                 "sample_weight, and additional keyword arguments.")
        assert isinstance(nepochs, int) and nepochs > 0
        if kwargs_cb is not None:
            assert isinstance(kwargs_cb, dict)
        else:
            kwargs_cb = dict()
        if random_state is None:
            random_state = 1

        want = np.float32 if use_float else np.float64
        if _is_torch(x0):
            import torch
            if x0.dtype != (torch.float32 if use_float else torch.float64):
                raise ValueError("'x0' has wrong dtype.")
            if x0.dim() > 1:
                raise ValueError("'x0' must be a 1-dimensional array.")
        else:
            if x0.dtype != want:
                raise ValueError("'x0' has wrong dtype.")
            if x0.ndim > 1:
                raise ValueError("'x0' must be a 1-dimensional array.")

        self.x = x0
        self.n = x0.shape[0]
        self.step_size = step_size
        self.obj_fun = obj_fun
        self.pred_fun = pred_fun
        self.grad_fun = grad_fun
        self.callback_epoch = callback_epoch
        self.callback_iter = callback_iter
        self.tol = tol
        self.nepochs = nepochs
        self.batches_per_epoch = batches_per_epoch
        self.decr_step_size = decr_step_size
        self.kwargs_cb = kwargs_cb
        self.valset_frac = valset_frac
        self.random_state = random_state
        self.use_float = bool(use_float)
        self.verbose = bool(verbose)
        self.shuffle_data = bool(shuffle_data)
        self.epoch = 0
        self.batch_size = None
        # the first call asks for the first gradient (no calculation needed yet)
        self.req = self.optimizer.run_optimizer(self.x, self.step_size)
        if self.optimizer_name != "oLBFGS":
            self.use_grad_diff = bool(use_grad_diff)
            self._stash = _RowStash()

    @property
    def niter(self):
        return self.optimizer.niter

    # ---- the long batch inside fit() (stochqn/_optimizers.py:55-78) -------------------------------------
    def _fit_long_rows(self, n_rows, batch):
        """Row range [first, last) of the long batch that ends with mini-batch number `batch`."""
        L = self.optimizer.bfgs_upd_freq
        diff = (batch + 1) % L
        span = L - diff
        if (batch + 1) >= span:
            first = (batch + 1 - span) * self.batch_size
            last = min(n_rows, (batch + 1) * self.batch_size)
        else:
            first, last = 0, min(n_rows, span * self.batch_size)
        return first, last, diff

    def _get_long_batch(self, X, y, w, batch):
        first, last, diff = self._fit_long_rows(X.shape[0], batch)
        X_long, y_long = X[first:last], y[first:last]
        w_long = w[first:last] if w is not None else None
        if diff > 0:
            # the schedule of pair updates is out of phase with the epochs: carry rows over in the stash
            self._stash.push(X_long, y_long, w_long)
            X_long, y_long, w_long = self._stash.pop_all()
        return X_long, y_long, w_long

    # ---- fit ----------------------------------------------------------------------------------------
    def fit(self, X, y, sample_weight=None, additional_kwargs={}, valset=None):
        """Fit to sample data in `batches_per_epoch` mini-batches per epoch for up to `nepochs` epochs
        (reference: stochqn/_optimizers.py:201-286)."""
        X, y, sample_weight, additional_kwargs = self._check_fit_inputs(X, y, sample_weight, additional_kwargs, check_sp=True)
        if valset is not None:
            if self.obj_fun is None:
                raise ValueError("Must provide objective function when using a validation set for monitoring.")
            assert isinstance(valset, tuple) and len(valset) == 3
            X_val, y_val, w_val = valset
            X_val, y_val, w_val, additional_kwargs = self._check_fit_inputs(X_val, y_val, w_val, additional_kwargs)
            if self.valset_frac is not None:
                warnings.warn("'valset_frac' is ignored when passing a validation set to '.fit'.")
        elif self.valset_frac is not None:
            X, X_val, y, y_val, sample_weight, w_val = self._split_validation(X, y, sample_weight)
        else:
            X_val, y_val, w_val = None, None, None

        obj_last_epoch = np.inf
        say_done = self.verbose
        n_rows = X.shape[0]
        self.batch_size = int(np.ceil(n_rows / self.batches_per_epoch))
        for self.epoch in range(self.nepochs):
            if self.shuffle_data:
                np.random.seed(self.random_state + self.epoch)
                order = np.argsort(np.random.random(size=n_rows))
                X, y, sample_weight = _take_rows(X, order), _take_rows(y, order), _take_rows(sample_weight, order)
            for batch in range(self.batches_per_epoch):
                r0 = batch * self.batch_size
                r1 = min(n_rows, (batch + 1) * self.batch_size)
                w_batch = sample_weight[r0:r1] if sample_weight is not None else None
                self._fit_batch(X[r0:r1], y[r0:r1], w_batch, additional_kwargs, is_user_batch=False,
                                X_full=X, y_full=y, w_full=sample_weight, X_val=X_val, y_val=y_val, w_val=w_val, batch=batch)
            if self.callback_epoch is not None:
                self.callback_epoch(self.x, **self.kwargs_cb)
            if X_val is not None and self.obj_fun is not None:
                obj_this_epoch = float(self.obj_fun(self.x, X_val, y_val, sample_weight=w_val, **additional_kwargs))
                if self.verbose:
                    print((self.optimizer_name + " - epoch: %2d, f(x): %12.4f") % (self.epoch + 1, obj_this_epoch))
                if (obj_last_epoch - obj_this_epoch) < self.tol and obj_this_epoch <= obj_last_epoch:
                    if self.verbose:
                        print(self.optimizer_name + " - Optimization procedure terminated (decrease below tolerance).")
                        say_done = False
                    break
                obj_last_epoch = obj_this_epoch
        if say_done:
            print(self.optimizer_name + " - Optimization procedure terminated (reached number of epochs).")
        return self

    def _split_validation(self, X, y, sample_weight):
        """The reference's hold-out split (sklearn ``train_test_split`` with `random_state`,
        stochqn/_optimizers.py:242-247), applied to row numbers so that device tensors are split on their device."""
        from sklearn.model_selection import train_test_split
        rows = np.arange(X.shape[0])
        tr, va = train_test_split(rows, test_size=self.valset_frac, random_state=self.random_state)
        w_tr = _take_rows(sample_weight, tr) if sample_weight is not None else None
        w_va = _take_rows(sample_weight, va) if sample_weight is not None else None
        return _take_rows(X, tr), _take_rows(X, va), _take_rows(y, tr), _take_rows(y, va), w_tr, w_va

    # ---- partial_fit --------------------------------------------------------------------------------
    def partial_fit(self, X, y, sample_weight=None, additional_kwargs={}):
        """Update the model with one user-provided batch (reference: stochqn/_optimizers.py:288-337).
        SQN / adaQN keep the batches since the last correction pair for the next long-batch request."""
        X, y, sample_weight, additional_kwargs = self._check_fit_inputs(X, y, sample_weight, additional_kwargs)
        keep = False
        if self.optimizer_name == "SQN":
            keep = True
        elif self.optimizer_name == "adaQN":
            max_incr = self.optimizer.max_incr
            keep = self.use_grad_diff or (max_incr is not None and max_incr > 0)
        if keep:
            self._stash.push(X, y, sample_weight)
        self._fit_batch(X, y, sample_weight, additional_kwargs, is_user_batch=True)
        return self

    # ---- the request loop on one mini-batch (stochqn/_optimizers.py:339-382) -----------------------------
    def _fit_batch(self, X_batch, y_batch, w_batch, additional_kwargs, is_user_batch=False,
                   X_full=None, y_full=None, w_full=None, X_val=None, y_val=None, w_val=None, batch=None):
        opt = self.optimizer
        while True:
            task = self.req["task"]
            at = self.req["requested_on"]
            if task in ("calc_grad", "calc_grad_same_batch"):
                opt.update_gradient(self.grad_fun(at, X_batch, y_batch, sample_weight=w_batch, **additional_kwargs))
            elif task == "calc_fun_val_batch" and X_val is not None:
                opt.update_function(self.obj_fun(at, X_val, y_val, sample_weight=w_val, **additional_kwargs))
            else:
                if is_user_batch:
                    X_long, y_long, w_long = self._stash.pop_all()
                else:
                    X_long, y_long, w_long = self._get_long_batch(X_full, y_full, w_full, batch)
                if task == "calc_grad_big_batch":
                    opt.update_gradient(self.grad_fun(at, X_long, y_long, sample_weight=w_long, **additional_kwargs))
                elif task == "calc_hess_vec":
                    opt.update_hess_vec(self.hess_vec_fun(at[0], at[1], X_long, y_long, sample_weight=w_long, **additional_kwargs))
                elif task == "calc_fun_val_batch":
                    opt.update_function(self.obj_fun(at, X_long, y_long, sample_weight=w_long, **additional_kwargs))
                else:
                    raise ValueError("Unexpected request from the optimizer: %r" % (task,))

            clock = self.niter if is_user_batch else self.epoch
            self.req = opt.run_optimizer(self.x, self.decr_step_size(self.step_size, clock))

            if self.verbose and self.req["info"]["iteration_info"] != "no_problems_encountered":
                where = (" - at iteration %3d: " % self.niter) if is_user_batch else \
                        (" - at iteration %3d, epoch %2d: " % (self.niter, self.epoch + 1))
                print(self.optimizer_name + where + self.req["info"]["iteration_info"])
            if self.req["task"] == "calc_grad":
                if self.callback_iter is not None:
                    self.callback_iter(self.x, **self.kwargs_cb)
                break

    # ---- predictions / access -------------------------------------------------------------------------
    def predict(self, X, additional_kwargs={}):
        """Predictions of the user-supplied `pred_fun` at the current variable values."""
        if self.pred_fun is None:
            raise ValueError("Must supply predict function in order to call this method.")
        return self.pred_fun(self.x, X, **additional_kwargs)

    def get_x(self):
        """A copy of the current variable values (same container type as `x0`)."""
        return self.x.clone() if _is_torch(self.x) else self.x.copy()


class oLBFGS(_StochQN):
    """oLBFGS optimizer, guided mode (reference: stochqn/_optimizers.py:416-522)."""

    def __init__(self, x0, grad_fun, obj_fun=None, pred_fun=None, batches_per_epoch=25, step_size=1e-3,
                 decr_step_size="auto", shuffle_data=True, random_state=1, nepochs=25, valset_frac=None, tol=1e-1,
                 callback_epoch=None, callback_iter=None, kwargs_cb={}, verbose=True,
                 mem_size=10, hess_init=None, min_curvature=1e-4, y_reg=None, check_nan=True, nthreads=-1, use_float=False):
        self.optimizer_name = "oLBFGS"
        self.optimizer = self._free_class()(mem_size, hess_init, min_curvature, y_reg, check_nan, nthreads, use_float)
        self._add_common_attributes(x0, batches_per_epoch, step_size, grad_fun, obj_fun, pred_fun, decr_step_size,
                                    callback_epoch, callback_iter, valset_frac, tol, nepochs, kwargs_cb, random_state,
                                    shuffle_data, verbose, True, use_float)

    @staticmethod
    def _free_class():
        return oLBFGS_free


class SQN(_StochQN):
    """SQN optimizer, guided mode (reference: stochqn/_optimizers.py:524-650)."""

    def __init__(self, x0, grad_fun, obj_fun=None, hess_vec_fun=None, pred_fun=None, batches_per_epoch=25, step_size=1e-3,
                 decr_step_size="auto", shuffle_data=True, random_state=1, nepochs=25, valset_frac=None, tol=1e-1,
                 callback_epoch=None, callback_iter=None, kwargs_cb={}, verbose=True,
                 mem_size=10, bfgs_upd_freq=20, min_curvature=1e-4, y_reg=None, use_grad_diff=False, check_nan=True,
                 nthreads=-1, use_float=False):
        if not use_grad_diff and hess_vec_fun is None:
            raise ValueError("If not using 'use_grad_diff', must provide function that evaluates Hessian-vector product.")
        if hess_vec_fun is not None:
            if use_grad_diff:
                # (the reference then passes use_grad_diff=None down, i.e. the Hessian-vector route is taken:
                #  stochqn/_optimizers.py:633-636 - reproduced)
                warnings.warn("Hessian-vector function is ignored when passing 'use_grad_diff=True'.")
                use_grad_diff = None
            elif not callable(hess_vec_fun):
                raise ValueError("'hess_vec_fun' must be a function that takes as input the values of 'x' and a "
                                 "vector, returning the Hessian-vector product.")
        self.optimizer_name = "SQN"
        self.optimizer = self._free_class()(mem_size, bfgs_upd_freq, min_curvature, y_reg, use_grad_diff, check_nan,
                                            nthreads, use_float)
        self._add_common_attributes(x0, batches_per_epoch, step_size, grad_fun, obj_fun, pred_fun, decr_step_size,
                                    callback_epoch, callback_iter, valset_frac, tol, nepochs, kwargs_cb, random_state,
                                    shuffle_data, verbose, use_grad_diff, use_float)
        self.hess_vec_fun = hess_vec_fun

    @staticmethod
    def _free_class():
        return SQN_free


class adaQN(_StochQN):
    """adaQN optimizer, guided mode (reference: stochqn/_optimizers.py:652-789)."""

    def __init__(self, x0, grad_fun, obj_fun=None, pred_fun=None, batches_per_epoch=25, step_size=1e-1,
                 decr_step_size=None, shuffle_data=True, random_state=1, nepochs=25, valset_frac=None, tol=1e-1,
                 callback_epoch=None, callback_iter=None, kwargs_cb={}, verbose=True,
                 mem_size=10, fisher_size=100, bfgs_upd_freq=20, max_incr=1.01, min_curvature=1e-4, y_reg=None,
                 scal_reg=1e-4, rmsprop_weight=None, use_grad_diff=False, check_nan=True, nthreads=-1, use_float=False):
        if max_incr is not None and obj_fun is None:
            raise ValueError("Must provide objective function when passing 'max_incr'.")
        if use_grad_diff and fisher_size is not None:
            warnings.warn("'fisher_size' ignored when using 'use_grad_diff=True'.")
        if fisher_size is None:
            use_grad_diff = True
        self.optimizer_name = "adaQN"
        self.optimizer = self._free_class()(mem_size, fisher_size, bfgs_upd_freq, max_incr, min_curvature, scal_reg,
                                            rmsprop_weight, y_reg, use_grad_diff, check_nan, nthreads, use_float)
        self._add_common_attributes(x0, batches_per_epoch, step_size, grad_fun, obj_fun, pred_fun, decr_step_size,
                                    callback_epoch, callback_iter, valset_frac, tol, nepochs, kwargs_cb, random_state,
                                    shuffle_data, verbose, use_grad_diff, use_float)

    @staticmethod
    def _free_class():
        return adaQN_free
