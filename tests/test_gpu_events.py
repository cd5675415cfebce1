"""GPU: forced events - the branches that decide the task sequence (SURVEY.md section 5 "failure detection"):
curvature rejection with full memory (quirk Q1), NaN / huge directions (stochqn.c:825-835), func_increased
(stochqn.c:1275-1283), check_nan = 0, invalid workspaces."""
import ctypes as C

import numpy as np
import pytest

from cuda_stepper import CudaStepper
from oracle import stochqn_np as O
from oracle.driver import HostStepper, discrete, run_trace
from oracle.problems import Logistic, Quadratic
from stochqn_b200 import _lib

pytestmark = pytest.mark.gpu
ORACLE = {"oLBFGS": O.OracleOLBFGS, "SQN": O.OracleSQN, "adaQN": O.OracleAdaQN}


@pytest.fixture(autouse=True, params=["three_launch", "one_launch"])
def step_route(request, monkeypatch):
    """Every forced event runs through both step routes of oLBFGS / SQN: K1 -> K2 -> K3 and the fused one-launch
    step used for latency-bound sizes (the library reads the threshold when a workspace is created)."""
    monkeypatch.setenv("STOCHQN_B200_SMALL_N", "0" if request.param == "three_launch" else str(1 << 30))


def _both(kind, kw, prob_f, calls, step, hooks_o, hooks_c, dtype=np.float64, mode="device"):
    p1, p2 = prob_f(), prob_f()
    so = HostStepper(ORACLE[kind](len(p1.x0()), dtype=dtype, **kw), p1.x0())
    sc = CudaStepper(kind, p2.x0(), dtype=dtype, mode=mode, **kw)
    to = run_trace(so, p1, calls, step, hooks=hooks_o, keep_x=True)
    tc = run_trace(sc, p2, calls, step, hooks=hooks_c, keep_x=True)
    return to, tc, so, sc


def _err(ta, tb):
    worst = 0.0
    for a, b in zip(ta, tb):
        fa, fb = np.isfinite(a["x"]), np.isfinite(b["x"])
        assert np.array_equal(fa, fb)
        if fb.any():
            worst = max(worst, float(np.max(np.abs(a["x"][fb] - b["x"][fb])) / max(np.max(np.abs(b["x"][fb])), 1e-300)))
    return worst


def test_rejected_pair_with_full_memory_zeroes_the_slot_q1():
    kw = dict(mem_size=2, hess_init=0.0, y_reg=0.0, min_curvature=1e-4, check_nan=1)
    prev = {}

    def remember(stepper, task, payload):       # call 7 = a step call: its incoming gradient becomes grad_prev
        prev[id(stepper)] = payload["grad"].copy()

    def zero_y(stepper, task, payload):          # call 8 = the pair call: hand back the same gradient -> y = 0
        payload["grad"] = prev[id(stepper)].copy()

    hooks = {7: remember, 8: zero_y}
    to, tc, so, sc = _both("oLBFGS", kw, lambda: Quadratic(6), 12, 1e-2, hooks, hooks)
    assert discrete(to) == discrete(tc)
    assert tc[8]["info"] == 202 and tc[8]["mem_used"] == 2 and tc[8]["mem_st_ix"] == 1
    assert tc[9]["info"] == 203 and tc[9]["ret"] == 0 and tc[9]["mem_used"] == 0 and tc[9]["niter"] == 5
    assert np.array_equal(tc[9]["x"], tc[8]["x"])
    assert _err(tc, to) <= 1e-10
    sc.close()


def test_slot_reads_zero_after_rejection():
    kw = dict(mem_size=2, hess_init=0.0, y_reg=0.0, min_curvature=1e-4, check_nan=1)
    q = Quadratic(6)
    sc = CudaStepper("oLBFGS", q.x0(), **kw)
    prev = {}
    hooks = {7: lambda s, t, p: prev.update(g=p["grad"].copy()), 8: lambda s, t, p: p.update(grad=prev["g"].copy())}
    run_trace(sc, q, 9, 1e-2, hooks=hooks)
    assert np.all(sc.slot("s", 1) == 0) and np.all(sc.slot("y", 1) == 0)
    assert np.any(sc.slot("s", 0) != 0)
    sc.close()


@pytest.mark.parametrize("kind,kw", [
    ("oLBFGS", dict(mem_size=3, hess_init=0.0, y_reg=0.0, min_curvature=1e-4, check_nan=1)),
    ("SQN", dict(mem_size=3, bfgs_upd_freq=3, min_curvature=1e-4, use_grad_diff=0, y_reg=0.0, check_nan=1)),
    ("adaQN", dict(mem_size=3, fisher_size=5, bfgs_upd_freq=3, max_incr=0.0, min_curvature=1e-4, scal_reg=1e-4,
                   rmsprop_weight=0.9, use_grad_diff=0, y_reg=0.0, check_nan=1)),
])
@pytest.mark.parametrize("poison", [np.nan, np.inf])
def test_nonfinite_gradient_rejects_step_and_flushes(kind, kw, poison):
    def bad(stepper, task, payload):
        if "grad" in payload:
            payload["grad"] = payload["grad"].copy()
            payload["grad"][2] = poison

    hooks = {13: bad}
    to, tc, so, sc = _both(kind, kw, lambda: Quadratic(6), 24, 5e-3, hooks, hooks)
    assert discrete(to) == discrete(tc)
    assert any(r["info"] == 203 for r in tc)
    assert _err(tc, to) <= 1e-9
    sc.close()


@pytest.mark.parametrize("kind,kw", [
    ("oLBFGS", dict(mem_size=3, hess_init=0.0, y_reg=0.0, min_curvature=1e-4, check_nan=1)),
    ("SQN", dict(mem_size=3, bfgs_upd_freq=3, min_curvature=1e-4, use_grad_diff=0, y_reg=0.0, check_nan=1)),
    ("SQN", dict(mem_size=3, bfgs_upd_freq=3, min_curvature=1e-4, use_grad_diff=1, y_reg=0.0, check_nan=1)),
])
@pytest.mark.parametrize("poison", [np.nan, np.inf])
def test_nonfinite_pair_in_slot_zero_is_survived(kind, kw, poison):
    """A non-finite same-batch / big-batch gradient (or Hessian-vector product) that lands in y slot 0 is accepted as a
    pair (NaN <= min_curvature is false, stochqn.c:892), the next step is rejected and the memory flushed, and the step
    after that runs with NO pairs: it must be the plain gradient step of stochqn.c:808-812 and must not read the
    poisoned slot.  Once with the empty memory at the start, once more right after the flush."""
    state = {"left": 2}

    def bad(stepper, task, payload):
        if task in (102, 103, 104) and state["left"] > 0 and stepper.counters()["mem_st_ix"] == 0 and stepper.counters()["mem_used"] == 0:
            key = "hess_vec" if task == 104 else "grad"
            if task == 103 and stepper.counters()["section"] == 2:
                return                          # first big-batch gradient of SQN only seeds grad_prev
            payload[key] = payload[key].copy()
            payload[key][1] = poison
            state["left"] -= 1

    traces = []
    for mk in ("oracle", "cuda"):
        state["left"] = 2
        p = Quadratic(6)
        st = HostStepper(ORACLE[kind](6, **kw), p.x0()) if mk == "oracle" else CudaStepper(kind, p.x0(), **kw)
        traces.append(run_trace(st, p, 40, 5e-3, hooks={c: bad for c in range(1, 40)}, keep_x=True))
        if mk == "cuda":
            st.close()
    to, tc = traces
    assert state["left"] == 0
    assert discrete(to) == discrete(tc)
    assert sum(r["info"] == 203 for r in tc) >= 2
    assert np.all(np.isfinite(tc[-1]["x"])) and np.all(np.isfinite(to[-1]["x"]))
    assert _err(tc, to) <= 1e-9


def test_check_nan_off_lets_nan_through_like_the_reference():
    kw = dict(mem_size=3, hess_init=0.0, y_reg=0.0, min_curvature=0.0, check_nan=0)

    def bad(stepper, task, payload):
        payload["grad"] = payload["grad"].copy()
        payload["grad"][2] = np.nan

    hooks = {9: bad}
    to, tc, so, sc = _both("oLBFGS", kw, lambda: Quadratic(6), 12, 1e-2, hooks, hooks)
    assert discrete(to) == discrete(tc)
    assert not np.all(np.isfinite(tc[9]["x"])) and not np.all(np.isfinite(to[9]["x"]))
    sc.close()


def test_huge_direction_takes_exact_norm_route_and_matches():
    """||d|| close to / beyond the reference's limit 1e3*n (stochqn.c:829): the cheap bound cannot certify the step,
    the library measures ||d|| exactly before touching x and decides exactly like the reference."""
    kw = dict(mem_size=3, hess_init=0.0, y_reg=0.0, min_curvature=1e-4, check_nan=1)

    def scale(f):
        def h(stepper, task, payload):
            payload["grad"] = payload["grad"] * f
        return h

    # limit = 6000 for n = 6; the quadratic's gradients are O(10): x 300 lands near the limit, x 3e5 far beyond it
    hooks = {7: scale(300.0), 11: scale(3.0e5)}
    to, tc, so, sc = _both("oLBFGS", kw, lambda: Quadratic(6), 16, 1e-4, hooks, hooks)
    assert discrete(to) == discrete(tc)
    assert _err(tc, to) <= 1e-9
    assert _lib.get_stat(sc.abi, sc.ws, _lib.STAT_EXACT_NORM_STEPS) >= 1
    assert any(r["info"] == 203 for r in tc)
    sc.close()


@pytest.mark.parametrize("gd", [0, 1])
def test_func_increased_reverts_x_and_flushes(gd):
    kw = dict(mem_size=5, fisher_size=20, bfgs_upd_freq=5, max_incr=1.01, min_curvature=1e-4, scal_reg=1e-4,
              rmsprop_weight=0.9, use_grad_diff=gd, y_reg=0.0, check_nan=1)

    def blow(stepper, task, payload):
        if "f" in payload:
            payload["f"] = 1e30

    hooks = {c: blow for c in range(22, 40)}
    for mode in ("device", "host"):
        to, tc, so, sc = _both("adaQN", kw, Logistic, 70, 1e-2, hooks, hooks, mode=mode)
        assert discrete(to) == discrete(tc)
        assert any(r["info"] == 201 for r in tc)
        assert _err(tc, to) <= 1e-10
        sc.close()


def test_invalid_workspace_is_answered_with_invalid_input():
    abi = _lib.load(np.float64)
    S = abi.structs
    fake = S["workspace_oLBFGS"]()          # a struct the caller assembled by hand: not ours
    fake.section = 1
    import torch
    x = torch.zeros(8, device="cuda", dtype=torch.float64)
    req, task, info = C.c_void_p(), C.c_int(), C.c_int()
    ret = abi.lib.run_oLBFGS(1e-3, x.data_ptr(), x.data_ptr(), C.byref(req), C.byref(task), C.byref(fake), C.byref(info))
    assert ret == -1000 and task.value == 100
    # a real workspace whose `section` was corrupted (stochqn.c:1033-1035)
    ws = abi.lib.initialize_oLBFGS(8, 3, 0.0, 0.0, 0.0, 1, 1)
    ws.contents.section = 7
    ret = abi.lib.run_oLBFGS(1e-3, x.data_ptr(), x.data_ptr(), C.byref(req), C.byref(task), ws, C.byref(info))
    assert ret == -1000 and task.value == 100
    abi.lib.dealloc_oLBFGS(ws)
    # parameters the GPU build cannot serve are refused at construction (NULL, like an allocation failure)
    assert not abi.lib.initialize_oLBFGS(8, 0, 0.0, 0.0, 0.0, 1, 1)
    assert not abi.lib.initialize_SQN(8, 1000, 3, 0.0, 0, 0.0, 1, 1)


def test_tunables_are_read_at_call_time():
    """y_reg, min_curvature, hess_init, check_nan, upd_freq may be changed between calls (include/stochqn.h:163-167)."""
    q1, q2 = Quadratic(6), Quadratic(6)
    kw = dict(mem_size=3, hess_init=0.0, y_reg=0.0, min_curvature=1e-6, check_nan=1)
    so = HostStepper(O.OracleOLBFGS(6, **kw), q1.x0())
    sc = CudaStepper("oLBFGS", q2.x0(), **kw)

    def tweak_o(stepper, task, payload):
        stepper.opt.hess_init = np.float64(0.02)
        stepper.opt.bfgs_memory.y_reg = np.float64(1e-2)
        stepper.opt.bfgs_memory.min_curvature = np.float64(1e-3)

    def tweak_c(stepper, task, payload):
        w = stepper.ws.contents
        w.hess_init = 0.02
        w.bfgs_memory.contents.y_reg = 1e-2
        w.bfgs_memory.contents.min_curvature = 1e-3

    to = run_trace(so, q1, 30, 1e-2, hooks={9: tweak_o}, keep_x=True)
    tc = run_trace(sc, q2, 30, 1e-2, hooks={9: tweak_c}, keep_x=True)
    assert discrete(to) == discrete(tc)
    assert _err(tc, to) <= 1e-10
    sc.close()
