"""CPU: the NumPy oracle against golden vectors produced by the reference C library
(tests/golden/traces_f64.json, generator tests/golden/make_golden.py) and against the
known answers recorded in SURVEY.md section 4."""
import json
import os

import numpy as np
import pytest

from cases import CASES, CASE_IDS
from oracle import stochqn_np as O
from oracle.driver import DISCRETE_FIELDS, HostStepper, run_trace
from oracle.problems import Quadratic, Rosenbrock

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "traces_f64.json")))
ORACLE = {"oLBFGS": O.OracleOLBFGS, "SQN": O.OracleSQN, "adaQN": O.OracleAdaQN}


@pytest.mark.parametrize("case", CASES, ids=CASE_IDS)
def test_oracle_reproduces_reference_trace(case):
    name, kind, kw, prob_f, calls, step = case
    gold = GOLD["cases"][name]
    p = prob_f()
    st = HostStepper(ORACLE[kind](len(p.x0()), dtype=np.float64, **kw), p.x0())
    tr = run_trace(st, p, calls, step, keep_x=True)
    mine = [[r.get(k) for k in DISCRETE_FIELDS] for r in tr]
    assert GOLD["fields"] == list(DISCRETE_FIELDS)
    for i, (a, b) in enumerate(zip(mine, gold["discrete"])):
        assert a == b, "call %d: oracle %r != reference %r" % (i, a, b)
    xn = np.array([r["x_norm"] for r in tr])
    assert np.allclose(xn, np.array(gold["x_norm"]), rtol=1e-7, atol=0)
    xf = np.array(gold["x_final"])
    # chaotic adaQN cases amplify last-bit differences between NumPy and OpenBLAS dots; 1e-6 still pins the algorithm
    assert np.max(np.abs(tr[-1]["x"] - xf)) <= 1e-6 * np.max(np.abs(xf))


def test_c_rosen_example_known_answer():
    """example/c_rosen.c output: f0 266.6000, f(10) 0.6755 ... f(200) 0.4916, final 0.4908,
    x = [1.048826 1.094597 1.253735 1.539749]."""
    prob = Rosenbrock(4, example_quirk=True)
    opt = O.OracleSQN(4, 5, 3, 0.0, 0, 1e-8, 1, 1)
    st = HostStepper(opt, np.array([1.3, 0.7, 0.8, 1.9]))
    assert "%.4f" % prob.fun(st.x) == "266.6000"
    ret, task, info = st.call(1e-3)
    printed = {}
    while opt.niter < 200:
        if task == 101:
            st.write("grad", prob.grad(st.read("req"), "new"))
        elif task == 104:
            st.write("hess_vec", prob.hess_vec(st.read("req"), st.read("req_vec")))
        ret, task, info = st.call(1e-3)
        if ret and (opt.niter + 1) % 10 == 0:
            printed[opt.niter + 1] = prob.fun(st.x)
    expect = {10: "0.6755", 20: "0.6633", 50: "0.6303", 100: "0.5798", 150: "0.5337", 200: "0.4916"}
    for k, v in expect.items():
        assert "%.4f" % printed[k] == v
    assert "%.4f" % prob.fun(st.x) == "0.4908"
    assert ["%.6f" % v for v in st.x] == ["1.048826", "1.094597", "1.253735", "1.539749"]
    gold = GOLD["c_rosen_example"]
    assert np.allclose(st.x, gold["x_final"], rtol=1e-12)


def _tasks(trace):
    return ["%d(%d,%d,%d,%d)" % (r["task"], r["ret"], r["niter"], r["mem_used"], r["mem_st_ix"]) for r in trace]


def test_task_sequences_of_survey_section4():
    """Golden request sequences on the 6-variable quadratic, m = 3, L = 3."""
    q = Quadratic(6)
    tr = run_trace(HostStepper(O.OracleOLBFGS(6, 3, 0.0, 0.0, 1e-4, 1, 1), q.x0()), Quadratic(6), 9, 1e-2)
    assert _tasks(tr) == ["101(0,0,0,0)", "102(1,1,0,0)", "101(0,1,1,1)", "102(1,2,1,1)", "101(0,2,2,2)", "102(1,3,2,2)",
                          "101(0,3,3,0)", "102(1,4,3,0)", "101(0,4,3,1)"]
    tr = run_trace(HostStepper(O.OracleSQN(6, 3, 3, 1e-4, 0, 0.0, 1, 1), q.x0()), Quadratic(6), 12, 1e-2)
    t = _tasks(tr)
    assert [s[:3] for s in t[:6]] == ["101"] * 6 and t[6] == "104(1,6,0,0)" and t[7] == "101(0,6,1,1)"
    assert tr[6]["req"] == "x_avg" and t[10] == "104(1,9,1,1)" and t[11] == "101(0,9,2,2)"
    tr = run_trace(HostStepper(O.OracleSQN(6, 3, 3, 1e-4, 1, 0.0, 1, 1), q.x0()), Quadratic(6), 12, 1e-2)
    t = _tasks(tr)
    assert t[3] == "103(1,3,0,0)" and tr[3]["req"] == "x_avg_prev" and t[4].startswith("101(0,3")
    assert t[7] == "103(1,6,0,0)" and tr[7]["req"] == "x_avg" and t[8] == "101(0,6,1,1)" and t[11] == "103(1,9,1,1)"
    mk = lambda gd, mi: O.OracleAdaQN(6, 3, 5, 3, mi, 1e-4, 1e-4, 0.9, gd, 0.0, 1, 1)   # noqa: E731
    tr = run_trace(HostStepper(mk(0, 0.0), q.x0()), Quadratic(6), 14, 5e-3)
    assert all(r["task"] == 101 for r in tr)
    assert [r["mem_used"] for r in tr if r["niter"] in (6, 9, 12)][0::1][0] == 1
    assert [r["fisher_used"] for r in tr][1:8] == [1, 2, 3, 4, 5, 5, 5]
    tr = run_trace(HostStepper(mk(0, 1.01), q.x0()), Quadratic(6), 13, 5e-3)
    t = _tasks(tr)
    assert t[3].startswith("105(1,3") and tr[3]["req"] == "x_avg_prev" and t[4].startswith("101(0,3")
    assert t[7].startswith("105(1,6") and tr[7]["req"] == "x_avg" and t[8] == "101(0,6,1,1)"
    tr = run_trace(HostStepper(mk(1, 1.01), q.x0()), Quadratic(6), 11, 5e-3)
    t = _tasks(tr)
    assert t[3].startswith("103(1,3") and t[4].startswith("105(0,3") and t[5].startswith("101(0,3")
    assert t[8].startswith("105(1,6") and t[9].startswith("103(0,6") and t[10] == "101(0,6,1,1)"


def test_rejection_with_full_memory_quirk_q1():
    """oLBFGS m=2, forced y = 0 at the 4th same-batch request: info 202, counters unchanged, the oldest slot
    zeroed; the NEXT call rejects the step (203), flushes the memory and leaves x alone."""
    q = Quadratic(6)
    opt = O.OracleOLBFGS(6, 2, 0.0, 0.0, 1e-4, 1, 1)
    st = HostStepper(opt, q.x0())
    same_batch_seen = {"n": 0}

    def force_zero_y(stepper, task, payload):
        payload["grad"] = stepper.opt.grad_prev.astype(np.float64).copy()      # y = grad - grad_prev = 0

    # calls: 0 init, then alternating step / pair.  The 4th same-batch request is answered at call 8.
    tr = run_trace(st, q, 10, 1e-2, hooks={8: force_zero_y}, keep_x=True)
    assert tr[8]["info"] == 202 and tr[8]["mem_used"] == 2 and tr[8]["mem_st_ix"] == 1
    assert tr[9]["info"] == 203 and tr[9]["ret"] == 0 and tr[9]["niter"] == 5 and tr[9]["mem_used"] == 0 and tr[9]["mem_st_ix"] == 0
    assert np.array_equal(tr[9]["x"], tr[8]["x"])
    del same_batch_seen
