"""Parity at BASELINE.json's full size (n = 2^27, mem_size 10, fp64) through a size-independent property.

The oracle cannot run n = 2^27 in seconds, but L-BFGS-type updates are *tiling invariant*: for a separable
objective made of c identical blocks of P variables, started from the same point in every block, every inner
product of the recursion is exactly c times the one-block value, so the coefficients rho_i * (s_i . q),
gamma = s.y / y.y, the curvature ratio s.y / s.s (stochqn.c:671-707, 883-900) and therefore the iterates are
those of the one-block problem repeated c times; AdaGrad / RMSProp scaling is element-wise, iterate averaging too.
(The empirical-Fisher product F'(F s)/k is NOT invariant - it scales with c - so adaQN is exercised in its
gradient-difference mode here and in Fisher mode against the oracle directly at small n, tests/test_gpu_parity.py.)

So: run the CUDA library on the tiled problem at full size with gradients evaluated on the device, run the
oracle on one block, and require identical task / return / counter / info sequences and iterates within 1e-10.
"""
from __future__ import annotations

import numpy as np
import pytest

from cuda_stepper import CudaStepper, _RawDeviceArray
from oracle import stochqn_np as O
from oracle.driver import (CALC_FUN_VAL_BATCH, CALC_GRAD, CALC_GRAD_BIG_BATCH, CALC_GRAD_SAME_BATCH, CALC_HESS_VEC,
                           HostStepper, run_trace)

pytestmark = pytest.mark.gpu

P = 1024


class Separable:
    """f(x) = sum_i x_i^4/4 + a_i x_i^2/2 - b_i x_i on one block of P variables (NumPy, fp64)."""

    def __init__(self, seed=5):
        rng = np.random.default_rng(seed)
        self.a = 0.5 + rng.random(P)
        self.b = rng.standard_normal(P)
        self.n = P

    def x0(self):
        return np.linspace(-1.5, 1.5, P)

    def grad(self, x, kind=None):
        return x ** 3 + self.a * x - self.b

    def hess_vec(self, x, v):
        return (3.0 * x * x + self.a) * v

    def fun(self, x):
        return float(np.sum(0.25 * x ** 4 + 0.5 * self.a * x * x - self.b * x)) + 2.0 * P      # kept positive (max_incr test)


def _run_tiled(kind, c, n_calls, step, kw):
    """The request loop on the c-times tiled problem, every callback evaluated by torch on the device."""
    import torch

    small = Separable()
    n = c * P
    a = torch.tensor(small.a, device="cuda")
    b = torch.tensor(small.b, device="cuda")
    st = CudaStepper(kind, np.zeros(n), **kw)              # np.zeros is lazily mapped: no 1 GiB host fill
    st.x.view(c, P).copy_(torch.tensor(small.x0(), device="cuda").unsqueeze(0).expand(c, P))

    def view(ptr):
        return torch.as_tensor(_RawDeviceArray(ptr, n, "<f8"), device="cuda").view(c, P)

    trace = []
    f = 0.0
    ret, task, info = st.call(step, 0.0)
    trace.append(dict(task=task, ret=ret, info=info, req=st.req_label, **st.counters()))
    for _ in range(1, n_calls):
        f = 0.0
        if task in (CALC_GRAD, CALC_GRAD_SAME_BATCH, CALC_GRAD_BIG_BATCH):
            xr = view(st._req.value)
            g = st.grad.view(c, P)
            torch.mul(xr, xr, out=g)
            g.mul_(xr).addcmul_(xr, a.unsqueeze(0)).sub_(b.unsqueeze(0))
        elif task == CALC_HESS_VEC:
            xr, vr = view(st._req.value), view(st._req_vec.value)
            h = st.hess_vec.view(c, P)
            torch.mul(xr, xr, out=h)
            h.mul_(3.0).add_(a.unsqueeze(0)).mul_(vr)
        elif task == CALC_FUN_VAL_BATCH:
            xr = view(st._req.value)[0]          # every block holds the same values: one block is enough
            f = (float((0.25 * xr ** 4 + 0.5 * a * xr * xr - b * xr).sum().item()) + 2.0 * P) * c
        else:
            raise RuntimeError("unexpected task %r" % task)
        ret, task, info = st.call(step, f)
        trace.append(dict(task=task, ret=ret, info=info, req=st.req_label, **st.counters()))
    torch.cuda.synchronize()
    xb = st.x.view(c, P)
    spread = float((xb - xb[0:1]).abs().max().item())
    x_block = xb[0].cpu().numpy().copy()
    x_last = xb[-1].cpu().numpy().copy()
    st.close()
    return trace, x_block, x_last, spread


FULL = [
    # the headline configuration: oLBFGS, n = 2^27, mem_size 10, gamma scaling, curvature threshold on
    ("olbfgs_n2p27", "oLBFGS", 1 << 27, 90, 2e-2,
     dict(mem_size=10, hess_init=0.0, y_reg=0.0, min_curvature=1e-4, check_nan=1)),
    ("sqn_hv_n2p26", "SQN", 1 << 26, 120, 2e-2,
     dict(mem_size=10, bfgs_upd_freq=3, min_curvature=1e-4, use_grad_diff=0, y_reg=0.0, check_nan=1)),
    ("adaqn_gd_n2p26", "adaQN", 1 << 26, 260, 1e-3,
     dict(mem_size=10, fisher_size=0, bfgs_upd_freq=10, max_incr=1.01, min_curvature=1e-4, scal_reg=1e-4,
          rmsprop_weight=0.0, use_grad_diff=1, y_reg=0.0, check_nan=1)),
]


@pytest.mark.parametrize("name,kind,n,calls,step,kw", FULL, ids=[f[0] for f in FULL])
def test_full_size_tiling_invariance(name, kind, n, calls, step, kw):
    import torch

    free, _total = torch.cuda.mem_get_info()
    need = (2 * kw["mem_size"] + 10) * n * 8
    if free < need:
        pytest.skip("needs %.0f GiB of device memory" % (need / 2 ** 30))
    c = n // P
    small = Separable()
    cls = {"oLBFGS": O.OracleOLBFGS, "SQN": O.OracleSQN, "adaQN": O.OracleAdaQN}[kind]
    so = HostStepper(cls(P, **kw), small.x0())
    to = run_trace(so, small, calls, step, keep_x=True)
    tc, x_block, x_last, spread = _run_tiled(kind, c, calls, step, kw)
    keys = ("task", "ret", "info", "req", "niter", "section", "mem_used", "mem_st_ix")
    assert [tuple(r[k] for k in keys) for r in tc] == [tuple(r[k] for k in keys) for r in to]
    assert to[-1]["mem_used"] == kw["mem_size"], "the case must run with the memory full"
    x_ref = to[-1]["x"]
    scale = np.max(np.abs(x_ref))
    assert np.max(np.abs(x_block - x_ref)) / scale <= 1e-10
    assert np.max(np.abs(x_last - x_ref)) / scale <= 1e-10
    assert spread / scale <= 1e-12, "blocks of the tiled problem drifted apart"
    assert np.linalg.norm(x_ref - small.x0()) > 1e-2, "the optimizer must have moved"
