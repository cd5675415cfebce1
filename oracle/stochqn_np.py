"""TEST INFRASTRUCTURE ONLY - CPU restatement of the reference optimizer step in NumPy.

This is the ORACLE the CUDA path is checked against.  It is never imported by the
product (``stochqn_b200/``); only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` leg of ``bench.py`` may use it, and only as the checker.

It restates, function by function, the algorithm of the reference core
``/root/reference/src/stochqn.c`` - the latency-chained two-loop recursion, NOT the
compact form the CUDA kernels use - so that agreement between the two is a real
check of the algebra.  Every function cites the reference lines it follows.
All quirks of SURVEY.md section 7.1 (Q1-Q9) are reproduced on purpose.

Pinned against: the reference C library itself, compiled unmodified into
``oracle/_ref`` (``oracle/build_ref.py``), through ``tests/test_oracle_vs_reference.py``
and the committed golden traces ``tests/golden/*.json`` (``tests/golden/make_golden.py``).

State layout mirrors the reference structs (include/stochqn.h:86-151) with NumPy
arrays in place of the C arrays; backup buffers are zero-filled (R semantics,
R/allocators.R:8-9 - the reference oracle is built with malloc->calloc to match).
"""
from __future__ import annotations

import numpy as np

# task_enum / info_enum values (include/stochqn.h:268-284)
CALC_GRAD, CALC_GRAD_SAME_BATCH, CALC_GRAD_BIG_BATCH, CALC_HESS_VEC, CALC_FUN_VAL_BATCH, INVALID_INPUT = \
    101, 102, 103, 104, 105, 100
NO_PROBLEMS, FUNC_INCREASED, CURVATURE_TOO_SMALL, SEARCH_DIRECTION_WAS_NAN = 200, 201, 202, 203


class BfgsMem:
    """bfgs_mem (include/stochqn.h:86-99; initialize_bfgs_mem stochqn.c:300-329)."""

    def __init__(self, mem_size, n, min_curvature, y_reg, upd_freq, dtype):
        self.s_mem = np.zeros((mem_size, n), dtype)
        self.y_mem = np.zeros((mem_size, n), dtype)
        self.buffer_rho = np.zeros(mem_size, dtype)
        self.buffer_alpha = np.zeros(mem_size, dtype)
        # zero-filled, never written (quirk Q1)
        self.s_bak = np.zeros(n, dtype) if min_curvature > 0 else None
        self.y_bak = np.zeros(n, dtype) if min_curvature > 0 else None
        self.mem_size = int(mem_size)
        self.mem_used = 0
        self.mem_st_ix = 0
        self.upd_freq = int(upd_freq)
        self.y_reg = dtype(y_reg)
        self.min_curvature = dtype(min_curvature)


class FisherMem:
    """fisher_mem (include/stochqn.h:101-107; initialize_fisher_mem stochqn.c:342-353)."""

    def __init__(self, mem_size, n, dtype):
        self.F = np.zeros((mem_size, n), dtype)
        self.buffer_y = np.zeros(mem_size, dtype)
        self.mem_size = int(mem_size)
        self.mem_used = 0
        self.mem_st_ix = 0


# ---- ring-buffer bookkeeping (stochqn.c:554-610) -------------------------------------------

def flush_bfgs_mem(m):                      # stochqn.c:554-558
    m.mem_used = 0
    m.mem_st_ix = 0


def flush_fisher_mem(f):                    # stochqn.c:560-567
    if f is not None:
        f.mem_used = 0
        f.mem_st_ix = 0


def incr_bfgs_counters(m):                  # stochqn.c:569-573
    m.mem_st_ix = (m.mem_st_ix + 1) % m.mem_size
    m.mem_used = min(m.mem_used + 1, m.mem_size)


def add_to_fisher_mem(grad, f):             # stochqn.c:581-587 (575-579 for the counters)
    if f is not None:
        f.F[f.mem_st_ix, :] = grad
        f.mem_st_ix = (f.mem_st_ix + 1) % f.mem_size
        f.mem_used = min(f.mem_used + 1, f.mem_size)


def backup_corr_pair(m):                    # stochqn.c:589-595 - copies FROM the backup (Q1)
    if m.min_curvature > 0:
        m.s_mem[m.mem_st_ix, :] = m.s_bak
        m.y_mem[m.mem_st_ix, :] = m.y_bak


def rollback_corr_pair(m):                  # stochqn.c:597-604
    if m.min_curvature > 0:
        m.s_mem[m.mem_st_ix, :] = m.s_bak
        m.y_mem[m.mem_st_ix, :] = m.y_bak
        return CURVATURE_TOO_SMALL
    return None


def archive_x_avg(x_avg, x_avg_prev):       # stochqn.c:606-610 (x_avg aliases x_sum, :134)
    x_avg_prev[:] = x_avg
    x_avg[:] = 0


def average_from_sum(arr_sum, n_summed):    # stochqn.c:286-291
    if n_summed > 1:
        arr_sum *= arr_sum.dtype.type(1) / arr_sum.dtype.type(n_summed)


# ---- the two-loop recursion (stochqn.c:663-708) ---------------------------------------------

def approx_inv_hess_grad(grad, H0, h0, m, mem_st_ix):
    """In place: grad <- H * grad.  `mem_st_ix` here is the OLDEST pair's slot, as passed by
    take_step (stochqn.c:820)."""
    dt = grad.dtype.type
    used, size = m.mem_used, m.mem_size
    rho, alpha = m.buffer_rho, m.buffer_alpha
    with np.errstate(all="ignore"):
        for ii in range(used):                                        # 671-679
            i = used - ii - 1
            ipos = (mem_st_ix + i) % size
            rho[i] = dt(1) / dt(np.dot(m.y_mem[ipos], m.s_mem[ipos]))
            alpha[i] = rho[i] * dt(np.dot(grad, m.s_mem[ipos]))
            grad -= alpha[i] * m.y_mem[ipos]
        if H0 is None and h0 <= 0:                                    # 683-689
            # size_t arithmetic: (mem_st_ix - 1 + mem_used) % mem_size with unsigned wrap (Q8)
            last_pos = ((mem_st_ix - 1 + used) % (1 << 64)) % size
            scaling = dt(np.dot(m.s_mem[last_pos], m.y_mem[last_pos])) / dt(np.dot(m.y_mem[last_pos], m.y_mem[last_pos]))
            grad *= scaling
        elif H0 is not None:                                          # 695
            grad *= H0
        else:                                                         # 698
            grad *= dt(h0)
        for i in range(used):                                         # 702-707
            ipos = (mem_st_ix + i) % size
            beta = rho[i] * dt(np.dot(m.y_mem[ipos], grad))
            grad += (alpha[i] - beta) * m.s_mem[ipos]


# ---- AdaGrad / RMSProp diagonal (stochqn.c:720-783) -----------------------------------------

def update_sum_sq(grad, grad_sum_sq, rmsprop_weight):                 # 720-747
    dt = grad.dtype.type
    if rmsprop_weight > 0 and rmsprop_weight < 1:
        w_new = dt(1) - dt(rmsprop_weight)
        grad_sum_sq[:] = dt(rmsprop_weight) * grad_sum_sq + w_new * (grad * grad)
    else:
        grad_sum_sq += grad * grad


def diag_rescal(direction, grad, grad_sum_sq, scal_reg, rmsprop_weight):   # 762-783
    update_sum_sq(grad, grad_sum_sq, rmsprop_weight)
    with np.errstate(all="ignore"):
        if direction is None:
            grad /= np.sqrt(grad_sum_sq + grad.dtype.type(scal_reg))
        else:
            # quirk Q2: the "H0" handed to the two-loop is the RESCALED GRADIENT
            direction[:] = grad / np.sqrt(grad_sum_sq + grad.dtype.type(scal_reg))


def check_inf_nan(arr):                                               # 228-266
    return not bool(np.all(np.isfinite(arr)))


def take_step(step_size, x, grad, m, rmsprop_weight, H0, h0, grad_sum_sq, scal_reg, check_nan):
    """stochqn.c:802-840.  Returns the info code it sets (or None)."""
    dt = grad.dtype.type
    n = x.shape[0]
    if m.mem_used == 0:                                               # 808-812
        if grad_sum_sq is not None:
            diag_rescal(None, grad, grad_sum_sq, scal_reg, rmsprop_weight)
    else:                                                             # 815-822
        if grad_sum_sq is not None:
            diag_rescal(H0, grad, grad_sum_sq, scal_reg, rmsprop_weight)
        oldest = 0 if m.mem_st_ix == m.mem_used else m.mem_st_ix      # 820 (Q8)
        approx_inv_hess_grad(grad, H0, h0, m, oldest)
    if check_nan:                                                     # 825-835
        with np.errstate(all="ignore"):
            bad = check_inf_nan(grad) or (np.sqrt(np.dot(grad.astype(np.float64), grad.astype(np.float64))) > 1e3 * n)
        if bad:
            flush_bfgs_mem(m)
            return SEARCH_DIRECTION_WAS_NAN
    with np.errstate(all="ignore"):
        x -= dt(step_size) * grad                                     # 838
    return None


# ---- correction pairs (stochqn.c:861-966) ----------------------------------------------------

def update_s_vector(x_sum, x_avg_prev, needs_div, m):                 # 861-870
    backup_corr_pair(m)
    if needs_div:
        average_from_sum(x_sum, m.upd_freq)
    m.s_mem[m.mem_st_ix, :] = x_sum - x_avg_prev


def check_min_curvature(m):                                           # 883-900
    """Returns CURVATURE_TOO_SMALL when the pair is rejected, else None (pair accepted)."""
    s = m.s_mem[m.mem_st_ix]
    y = m.y_mem[m.mem_st_ix]
    dt = s.dtype.type
    if m.min_curvature > 0:
        with np.errstate(all="ignore"):
            curv = dt(np.dot(s, y)) / dt(np.dot(s, s))
        if curv <= m.min_curvature:
            return rollback_corr_pair(m)
    incr_bfgs_counters(m)
    return None


def update_y_grad_diff(grad, grad_prev, m):                           # 915-926
    s = m.s_mem[m.mem_st_ix]
    y = m.y_mem[m.mem_st_ix]
    with np.errstate(all="ignore"):
        y[:] = grad - grad_prev
        if m.y_reg > 0:
            y += m.y_reg * s
    return check_min_curvature(m)


def update_y_fisher(f, m):                                            # 936-952
    s = m.s_mem[m.mem_st_ix]
    y = m.y_mem[m.mem_st_ix]
    dt = s.dtype.type
    k = f.mem_used
    with np.errstate(all="ignore"):
        f.buffer_y[:k] = f.F[:k] @ s                                  # 946-947: first k PHYSICAL rows
        y[:] = (dt(1) / dt(k)) * (f.F[:k].T @ f.buffer_y[:k])         # 948-949
    return check_min_curvature(m)


def update_y_hessvec(hess_vec, m):                                    # 962-966
    m.y_mem[m.mem_st_ix, :] = hess_vec
    return check_min_curvature(m)


# ---- the three state machines (stochqn.c:978-1315) -------------------------------------------

class OracleOLBFGS:
    """workspace_oLBFGS + run_oLBFGS (include/stochqn.h:109-118; stochqn.c:464-481, 978-1036)."""
    kind = "oLBFGS"

    def __init__(self, n, mem_size=10, hess_init=0.0, y_reg=0.0, min_curvature=0.0, check_nan=1,
                 nthreads=1, dtype=np.float64):
        self.dtype = np.dtype(dtype).type
        self.bfgs_memory = BfgsMem(mem_size, n, min_curvature, y_reg, 1, self.dtype)
        self.grad_prev = np.zeros(n, self.dtype)
        self.hess_init = self.dtype(hess_init)
        self.niter = 0
        self.section = 0
        self.check_nan = int(check_nan)
        self.n = int(n)
        self.req = None
        self.req_label = None

    def run(self, step_size, x, grad):
        """Returns (ret, task, info); the request point is self.req (label self.req_label)."""
        m = self.bfgs_memory
        info = NO_PROBLEMS
        if self.section == 0:                                         # 983-989
            self.section = 1
            self.req, self.req_label = x, "x"
            return 0, CALC_GRAD, info
        if self.section == 1:                                         # 992-1021
            self.grad_prev[:] = grad
            bad = take_step(step_size, x, grad, m, 0, None, self.hess_init, None, 0, self.check_nan)
            if bad is not None:
                info = bad
            self.niter += 1                                           # Q7: even when rejected
            self.req, self.req_label = x, "x"
            if info == NO_PROBLEMS:
                backup_corr_pair(m)
                grad *= -self.dtype(step_size)
                m.s_mem[m.mem_st_ix, :] = grad
                self.section = 2
                return 1, CALC_GRAD_SAME_BATCH, info
            flush_bfgs_mem(m)
            self.section = 1
            return 0, CALC_GRAD, info
        if self.section == 2:                                         # 1024-1031
            r = update_y_grad_diff(grad, self.grad_prev, m)
            if r is not None:
                info = r
            self.section = 1
            self.req, self.req_label = x, "x"
            return 0, CALC_GRAD, info
        return -1000, INVALID_INPUT, info                             # 1033-1035


class OracleSQN:
    """workspace_SQN + run_SQN (include/stochqn.h:120-131; stochqn.c:483-506, 1038-1153)."""
    kind = "SQN"

    def __init__(self, n, mem_size=10, bfgs_upd_freq=10, min_curvature=1e-4, use_grad_diff=0, y_reg=0.0,
                 check_nan=1, nthreads=1, dtype=np.float64):
        self.dtype = np.dtype(dtype).type
        self.bfgs_memory = BfgsMem(mem_size, n, min_curvature, y_reg, bfgs_upd_freq, self.dtype)
        self.grad_prev = np.zeros(n, self.dtype) if use_grad_diff else None
        self.x_sum = np.zeros(n, self.dtype)
        self.x_avg_prev = np.zeros(n, self.dtype)
        self.use_grad_diff = int(use_grad_diff)
        self.niter = 0
        self.section = 0
        self.check_nan = int(check_nan)
        self.n = int(n)
        self.req = self.req_vec = None
        self.req_label = None

    def _resume(self, x, ret, info):                                  # 1148-1152
        self.section = 1
        self.req, self.req_label = x, "x"
        return ret, CALC_GRAD, info

    def run(self, step_size, x, grad, hess_vec=None):
        m = self.bfgs_memory
        info = NO_PROBLEMS
        ret = 0
        if self.section == 0:                                         # 1044-1048
            return self._resume(x, ret, info)
        if self.section == 1:                                         # 1051-1115
            bad = take_step(step_size, x, grad, m, 0, None, 0, None, 0, self.check_nan)
            if bad is not None:
                info = bad
            self.niter += 1
            ret = 0 if info == SEARCH_DIRECTION_WAS_NAN else 1
            self.x_sum += x                                           # 1067 (Q7)
            if self.niter % m.upd_freq != 0:
                return self._resume(x, ret, info)
            if self.niter == m.upd_freq:                              # 1078-1094
                average_from_sum(self.x_sum, m.upd_freq)
                archive_x_avg(self.x_sum, self.x_avg_prev)
                if self.use_grad_diff:
                    self.section = 2
                    self.req, self.req_label = self.x_avg_prev, "x_avg_prev"
                    return ret, CALC_GRAD_BIG_BATCH, info
                return self._resume(x, ret, info)
            update_s_vector(self.x_sum, self.x_avg_prev, 1, m)        # 1097
            self.req, self.req_label = self.x_sum, "x_avg"
            if self.use_grad_diff:                                    # 1100-1105
                self.section = 3
                return ret, CALC_GRAD_BIG_BATCH, info
            self.section = 4                                          # 1107-1113
            self.req_vec = m.s_mem[m.mem_st_ix]
            return ret, CALC_HESS_VEC, info
        if self.section == 2:                                         # 1118-1122
            self.grad_prev[:] = grad
            return self._resume(x, ret, info)
        if self.section == 3:                                         # 1125-1134
            r = update_y_grad_diff(grad, self.grad_prev, m)
            if r is not None:
                info = r
            if info == NO_PROBLEMS:
                self.grad_prev[:] = grad
                self.x_avg_prev[:] = self.x_sum
            self.x_sum[:] = 0
            return self._resume(x, ret, info)
        if self.section == 4:                                         # 1137-1142 (Q6)
            archive_x_avg(self.x_sum, self.x_avg_prev)
            r = update_y_hessvec(hess_vec, m)
            if r is not None:
                info = r
            return self._resume(x, ret, info)
        return -1000, INVALID_INPUT, info                             # 1144-1146


class OracleAdaQN:
    """workspace_adaQN + run_adaQN (include/stochqn.h:133-151; stochqn.c:508-547, 1155-1315)."""
    kind = "adaQN"

    def __init__(self, n, mem_size=10, fisher_size=100, bfgs_upd_freq=10, max_incr=1.01, min_curvature=1e-4,
                 scal_reg=1e-4, rmsprop_weight=0.9, use_grad_diff=0, y_reg=0.0, check_nan=1, nthreads=1,
                 dtype=np.float64):
        self.dtype = np.dtype(dtype).type
        self.bfgs_memory = BfgsMem(mem_size, n, min_curvature, y_reg, bfgs_upd_freq, self.dtype)
        if use_grad_diff:                                             # 515-521
            self.fisher_memory = None
            self.grad_prev = np.zeros(n, self.dtype)
        else:
            self.fisher_memory = FisherMem(fisher_size, n, self.dtype)
            self.grad_prev = None
        self.H0 = np.zeros(n, self.dtype)
        self.x_sum = np.zeros(n, self.dtype)
        self.x_avg_prev = np.zeros(n, self.dtype)
        self.grad_sum_sq = np.zeros(n, self.dtype)
        self.max_incr = self.dtype(max_incr)
        self.scal_reg = self.dtype(scal_reg)
        self.rmsprop_weight = self.dtype(rmsprop_weight)
        self.use_grad_diff = int(use_grad_diff)
        self.f_prev = self.dtype(0)
        self.niter = 0
        self.section = 0
        self.check_nan = int(check_nan)
        self.n = int(n)
        self.req = None
        self.req_label = None

    def _resume(self, x, ret, info):                                  # 1310-1314
        self.section = 1
        self.req, self.req_label = x, "x"
        return ret, CALC_GRAD, info

    def _update_y(self, x, ret, info):                                # 1297-1308
        m = self.bfgs_memory
        if self.use_grad_diff:
            self.req, self.req_label = self.x_sum, "x_avg"
            self.section = 4
            return ret, CALC_GRAD_BIG_BATCH, info
        r = update_y_fisher(self.fisher_memory, m)
        if r is not None:
            info = r
        if info == NO_PROBLEMS:
            self.x_avg_prev[:] = self.x_sum
        self.x_sum[:] = 0
        return self._resume(x, ret, info)

    def run(self, step_size, x, f, grad):
        m = self.bfgs_memory
        info = NO_PROBLEMS
        ret = 0
        if self.section == 0:                                         # 1161-1165
            return self._resume(x, ret, info)
        if self.section == 1:                                         # 1170-1239
            add_to_fisher_mem(grad, self.fisher_memory)               # 1174: raw gradient
            bad = take_step(step_size, x, grad, m, self.rmsprop_weight, self.H0, 0, self.grad_sum_sq,
                            self.scal_reg, self.check_nan)
            if bad is not None:
                info = bad
            ret = 0 if info == SEARCH_DIRECTION_WAS_NAN else 1
            self.niter += 1
            self.x_sum += x                                           # 1191
            if self.niter % m.upd_freq != 0:
                return self._resume(x, ret, info)
            if self.niter == m.upd_freq:                              # 1206-1224
                average_from_sum(self.x_sum, m.upd_freq)
                archive_x_avg(self.x_sum, self.x_avg_prev)
                if self.use_grad_diff:
                    self.req, self.req_label = self.x_avg_prev, "x_avg_prev"
                    self.section = 2
                    return ret, CALC_GRAD_BIG_BATCH, info
                if self.max_incr > 0:
                    self.req, self.req_label = self.x_avg_prev, "x_avg_prev"
                    self.section = 3
                    return ret, CALC_FUN_VAL_BATCH, info
                return self._resume(x, ret, info)
            if self.max_incr > 0:                                     # 1227-1234
                average_from_sum(self.x_sum, m.upd_freq)
                self.req, self.req_label = self.x_sum, "x_avg"
                self.section = 5
                return ret, CALC_FUN_VAL_BATCH, info
            update_s_vector(self.x_sum, self.x_avg_prev, 1, m)        # 1237
            return self._update_y(x, ret, info)
        if self.section == 2:                                         # 1242-1255
            self.grad_prev[:] = grad
            if self.max_incr:
                self.req, self.req_label = self.x_avg_prev, "x_avg_prev"
                self.section = 3
                return 0, CALC_FUN_VAL_BATCH, info
            return self._resume(x, ret, info)
        if self.section == 3:                                         # 1258-1262
            self.f_prev = self.dtype(f)
            return self._resume(x, ret, info)
        if self.section == 4:                                         # 1265-1270 (Q4)
            r = update_y_grad_diff(grad, self.grad_prev, m)
            if r is not None:
                info = r
            if info == NO_PROBLEMS:
                self.grad_prev[:] = grad
            self.x_sum[:] = 0
            return self._resume(x, ret, info)
        if self.section == 5:                                         # 1273-1291
            f = self.dtype(f)
            if f > self.max_incr * self.f_prev or np.isinf(f) or np.isnan(f):
                flush_bfgs_mem(m)
                flush_fisher_mem(self.fisher_memory)
                x[:] = self.x_avg_prev
                info = FUNC_INCREASED                                  # Q5: x_sum is not reset
                return self._resume(x, 1, info)
            self.f_prev = f
            update_s_vector(self.x_sum, self.x_avg_prev, 0, m)
            return self._update_y(x, ret, info)
        return -1000, INVALID_INPUT, info                             # 1293-1295
