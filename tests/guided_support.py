"""TEST INFRASTRUCTURE ONLY - shared pieces of the guided-mode parity tests.

* seeded binary-logistic data and NumPy callbacks with the reference's guided-mode signatures
  (``grad_fun(x, X, y, sample_weight=None, **kw)``, ``hess_vec_fun(x, v, X, y, ...)``, ``obj_fun``), closed forms of
  R/logistic.R:1-37;
* free-mode optimizer classes backed by the oracle (NumPy restatement) or by ``oracle/_ref`` (the reference C
  library), with the constructor / ``run_optimizer`` interface of stochqn/_optimizers.py:882-1364, so that the
  guided layer of this repository can be driven on CPU by something that is *not* the CUDA library;
* the case matrix shared by the golden generator (tests/golden/make_golden_guided.py, which runs the REFERENCE's
  own guided classes), the CPU tests and the GPU tests.
"""
from __future__ import annotations

import numpy as np

from oracle import stochqn_np as O

TASKS = {101: "calc_grad", 102: "calc_grad_same_batch", 103: "calc_grad_big_batch", 104: "calc_hess_vec",
         105: "calc_fun_val_batch"}
INFOS = {200: "no_problems_encountered", 201: "func_increased", 202: "curvature_too_small",
         203: "search_direction_was_nan"}


# ---- data and callbacks ---------------------------------------------------------------------------------
def make_data(nrows=1200, ncols=12, seed=7, weights=False):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((nrows, ncols))
    X[:, 0] = 1.0
    w_true = rng.standard_normal(ncols)
    y = (rng.random(nrows) < 1.0 / (1.0 + np.exp(-X @ w_true))).astype(np.float64)
    sw = (0.5 + rng.random(nrows)) if weights else None
    return X, y, sw


def _wts(X, sample_weight):
    return np.ones(X.shape[0]) if sample_weight is None else np.asarray(sample_weight, dtype=np.float64).reshape(-1)


def grad_fun(w, X, y, sample_weight=None, reg_param=0.0):
    sw = _wts(X, sample_weight)
    p = 1.0 / (1.0 + np.exp(-(X @ w)))
    return X.T @ ((p - y) * sw) / sw.sum() + 2.0 * reg_param * w


def hess_vec_fun(w, v, X, y, sample_weight=None, reg_param=0.0):
    sw = _wts(X, sample_weight)
    p = 1.0 / (1.0 + np.exp(-(X @ w)))
    return X.T @ (p * (1.0 - p) * sw * (X @ v)) / sw.sum() + 2.0 * reg_param * v


def obj_fun(w, X, y, sample_weight=None, reg_param=0.0):
    sw = _wts(X, sample_weight)
    z = X @ w
    ll = np.logaddexp(0.0, z) - y * z
    return float((ll * sw).sum() / sw.sum() + reg_param * (w @ w))


def pred_fun(w, X):
    return 1.0 / (1.0 + np.exp(-(X @ w)))


# ---- free-mode classes backed by the oracle / the reference library ------------------------------------------
class _HostFree:
    """run_optimizer / update_* of the free-mode interface over an in-place host optimizer object."""

    backend = "oracle"

    def _make(self, kind, n, **kw):
        if self.backend == "oracle":
            cls = {"oLBFGS": O.OracleOLBFGS, "SQN": O.OracleSQN, "adaQN": O.OracleAdaQN}[kind]
            return cls(n, **kw)
        from oracle import ref_lib
        cls = {"oLBFGS": ref_lib.RefOLBFGS, "SQN": ref_lib.RefSQN, "adaQN": ref_lib.RefAdaQN}[kind]
        return cls(n, **kw)

    def _common(self, min_curvature, y_reg, check_nan):
        self.min_curvature = 0.0 if min_curvature is None else min_curvature
        self.y_reg = 0.0 if y_reg is None else y_reg
        self.check_nan = int(bool(check_nan))
        self.initialized = False
        self.trace = []

    @property
    def niter(self):
        return int(self._opt.niter) if self.initialized else 0

    def update_gradient(self, g):
        self.gradient[:] = np.asarray(g, dtype=np.float64).reshape(-1)

    def _request(self, ret, task, info, at):
        self.trace.append((int(task), int(info), int(ret)))
        return {"task": TASKS[task], "requested_on": at,
                "info": {"x_changed_in_run": bool(ret), "iteration_number": self.niter, "iteration_info": INFOS[info]}}


class OracleOLBFGSFree(_HostFree):
    def __init__(self, mem_size=10, hess_init=None, min_curvature=1e-4, y_reg=None, check_nan=True, nthreads=-1,
                 use_float=False):
        assert not use_float
        self._common(min_curvature, y_reg, check_nan)
        self.mem_size = mem_size
        self.hess_init = 0.0 if hess_init is None else hess_init

    def run_optimizer(self, x, step_size):
        if not self.initialized:
            n = x.shape[0]
            self._opt = self._make("oLBFGS", n, mem_size=self.mem_size, hess_init=self.hess_init, y_reg=self.y_reg,
                                   min_curvature=self.min_curvature, check_nan=self.check_nan)
            self.gradient = np.zeros(n)
            self.initialized = True
        ret, task, info = self._opt.run(step_size, x, self.gradient)
        return self._request(ret, task, info, self._opt.req)


class OracleSQNFree(_HostFree):
    def __init__(self, mem_size=10, bfgs_upd_freq=20, min_curvature=1e-4, y_reg=None, use_grad_diff=False,
                 check_nan=True, nthreads=-1, use_float=False):
        assert not use_float
        self._common(min_curvature, y_reg, check_nan)
        self.mem_size = mem_size
        self.bfgs_upd_freq = int(bfgs_upd_freq)
        self.use_grad_diff = bool(use_grad_diff)

    def update_hess_vec(self, hv):
        self.hess_vec[:] = np.asarray(hv, dtype=np.float64).reshape(-1)

    def run_optimizer(self, x, step_size):
        if not self.initialized:
            n = x.shape[0]
            self._opt = self._make("SQN", n, mem_size=self.mem_size, bfgs_upd_freq=self.bfgs_upd_freq,
                                   min_curvature=self.min_curvature, use_grad_diff=int(self.use_grad_diff),
                                   y_reg=self.y_reg, check_nan=self.check_nan)
            self.gradient = np.zeros(n)
            self.hess_vec = np.zeros(n)
            self.initialized = True
        ret, task, info = self._opt.run(step_size, x, self.gradient, self.hess_vec)
        at = (self._opt.req, self._opt.req_vec) if task == 104 else self._opt.req
        return self._request(ret, task, info, at)


class OracleAdaQNFree(_HostFree):
    def __init__(self, mem_size=10, fisher_size=100, bfgs_upd_freq=20, max_incr=1.01, min_curvature=1e-4, scal_reg=1e-4,
                 rmsprop_weight=0.9, y_reg=None, use_grad_diff=False, check_nan=True, nthreads=-1, use_float=False):
        assert not use_float
        self._common(min_curvature, y_reg, check_nan)
        self.mem_size = mem_size
        self.use_grad_diff = bool(use_grad_diff)
        self.fisher_size = 0 if self.use_grad_diff else int(fisher_size)
        self.bfgs_upd_freq = int(bfgs_upd_freq)
        self.max_incr = 0.0 if max_incr is None else max_incr
        self.scal_reg = scal_reg
        self.rmsprop_weight = 0.0 if rmsprop_weight is None else rmsprop_weight
        self.f = 0.0

    def update_function(self, f):
        self.f = float(f)

    def run_optimizer(self, x, step_size):
        if not self.initialized:
            n = x.shape[0]
            self._opt = self._make("adaQN", n, mem_size=self.mem_size, fisher_size=self.fisher_size,
                                   bfgs_upd_freq=self.bfgs_upd_freq, max_incr=self.max_incr,
                                   min_curvature=self.min_curvature, scal_reg=self.scal_reg,
                                   rmsprop_weight=self.rmsprop_weight, use_grad_diff=int(self.use_grad_diff),
                                   y_reg=self.y_reg, check_nan=self.check_nan)
            self.gradient = np.zeros(n)
            self.initialized = True
        ret, task, info = self._opt.run(step_size, x, self.f, self.gradient)
        return self._request(ret, task, info, self._opt.req)


ORACLE_FREE = {"oLBFGS": OracleOLBFGSFree, "SQN": OracleSQNFree, "adaQN": OracleAdaQNFree}


# ---- the case matrix ---------------------------------------------------------------------------------------
# (name, optimizer, constructor kwargs, how it is driven)
#   mode "fit": one call to fit(X, y, sw);  mode "partial": partial_fit over the listed row ranges
REG = dict(reg_param=1e-3)
GUIDED_CASES = [
    ("olbfgs_fit_shuffle", "oLBFGS",
     dict(batches_per_epoch=10, step_size=2e-1, decr_step_size="auto", shuffle_data=True, random_state=3, nepochs=3,
          mem_size=5, min_curvature=1e-4), dict(mode="fit", weights=False)),
    ("olbfgs_partial", "oLBFGS",
     dict(batches_per_epoch=10, step_size=2e-1, decr_step_size=None, mem_size=4, hess_init=0.5, min_curvature=None, y_reg=1e-3),
     dict(mode="partial", weights=False, ranges=[(0, 100), (100, 250), (250, 300), (300, 520), (520, 640), (640, 900)])),
    ("sqn_hv_fit_phase", "SQN",
     dict(batches_per_epoch=10, step_size=2e-1, decr_step_size="auto", shuffle_data=False, nepochs=3,
          mem_size=5, bfgs_upd_freq=4, min_curvature=1e-4), dict(mode="fit", weights=True, hess_vec=True)),
    ("sqn_gd_fit_shuffle", "SQN",
     dict(batches_per_epoch=12, step_size=1e-1, decr_step_size=None, shuffle_data=True, random_state=11, nepochs=2,
          mem_size=5, bfgs_upd_freq=3, min_curvature=1e-4, use_grad_diff=True), dict(mode="fit", weights=False)),
    ("sqn_hv_partial", "SQN",
     dict(batches_per_epoch=10, step_size=1e-1, decr_step_size=None, mem_size=4, bfgs_upd_freq=3, min_curvature=1e-4),
     dict(mode="partial", weights=True, hess_vec=True,
          ranges=[(0, 90), (90, 200), (200, 260), (500, 640), (260, 380), (380, 500), (640, 800), (800, 900), (900, 1000),
                  (1000, 1100), (1100, 1200), (0, 150)])),
    ("adaqn_fisher_fit_valset", "adaQN",
     dict(batches_per_epoch=10, step_size=5e-2, decr_step_size=None, shuffle_data=True, random_state=5, nepochs=6,
          valset_frac=0.2, tol=1e-3, mem_size=5, fisher_size=12, bfgs_upd_freq=4, max_incr=1.01, min_curvature=1e-4,
          rmsprop_weight=0.9), dict(mode="fit", weights=False, obj=True)),
    ("adaqn_gd_partial", "adaQN",
     # (with max_incr set the reference's partial_fit fails: the function value and the big-batch gradient are two
     #  long-batch requests in a row and the second finds the stash empty - see test_guided_host.py)
     dict(batches_per_epoch=10, step_size=5e-2, decr_step_size=None, mem_size=5, fisher_size=None, bfgs_upd_freq=3,
          max_incr=None, min_curvature=1e-4, rmsprop_weight=None, use_grad_diff=True),
     dict(mode="partial", weights=False,
          ranges=[(0, 100), (100, 200), (200, 300), (300, 400), (400, 500), (500, 600), (600, 700), (700, 800),
                  (800, 900), (900, 1000), (1000, 1100), (1100, 1200)])),
    # CSR model matrix (the reference's fit accepts scipy CSR, stochqn/_optimizers.py:47-53): batches are CSR row slices, the
    # out-of-phase long batch goes through the stash and scipy.sparse.vstack; callbacks see sparse X
    ("sqn_hv_fit_csr", "SQN",
     dict(batches_per_epoch=9, step_size=2e-1, decr_step_size="auto", shuffle_data=True, random_state=2, nepochs=2,
          mem_size=4, bfgs_upd_freq=4, min_curvature=1e-4), dict(mode="fit", weights=True, hess_vec=True, sparse=True)),
    ("adaqn_fisher_fit_nomax", "adaQN",
     dict(batches_per_epoch=8, step_size=5e-2, decr_step_size="auto", shuffle_data=False, nepochs=3,
          mem_size=4, fisher_size=10, bfgs_upd_freq=3, max_incr=None, min_curvature=1e-4, rmsprop_weight=None),
     dict(mode="fit", weights=True)),
]
GUIDED_IDS = [c[0] for c in GUIDED_CASES]


def drive(cls, name, okw, how, to_array=lambda a: a, callbacks=None):
    """Construct guided class `cls` for one case and run it; returns (object, per-epoch/batch snapshots of x).

    `to_array` converts the NumPy data to whatever container the implementation under test wants
    (identity, or host -> torch CUDA tensor); `callbacks` = (grad, hess_vec, obj) override the NumPy ones."""
    X, y, sw = make_data(weights=how.get("weights", False))
    if how.get("sparse"):
        from scipy.sparse import csr_matrix
        X[np.abs(X) < 0.8] = 0.0                      # make it worth storing sparsely (deterministic: same seeded data)
        X[:, 0] = 1.0
        X = csr_matrix(X)
    g, hv, ob = callbacks or (grad_fun, hess_vec_fun, obj_fun)
    kw = dict(okw)
    kw["verbose"] = False
    snaps = []

    def snap(x, **_):
        snaps.append(np.array(x.detach().cpu().numpy() if hasattr(x, "detach") else x, dtype=np.float64))

    x0 = to_array(np.zeros(X.shape[1]))
    args = dict(x0=x0, grad_fun=g, pred_fun=None)
    if how.get("obj") or kw.get("valset_frac") is not None:
        args["obj_fun"] = ob
    if how.get("hess_vec"):
        args["hess_vec_fun"] = hv
    if how["mode"] == "fit":
        obj = cls(callback_epoch=snap, **args, **kw)
        obj.fit(X if how.get("sparse") else to_array(X), to_array(y), to_array(sw) if sw is not None else None,
                additional_kwargs=dict(REG))
    else:
        obj = cls(callback_iter=snap, **args, **kw)
        Xa, ya = to_array(X), to_array(y)
        swa = to_array(sw) if sw is not None else None
        for r0, r1 in how["ranges"]:
            obj.partial_fit(Xa[r0:r1], ya[r0:r1], swa[r0:r1] if swa is not None else None, additional_kwargs=dict(REG))
    return obj, snaps
