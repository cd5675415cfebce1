"""GPU: the Python free-mode classes (mirror of the reference's stochqn/_optimizers.py:882-1364) on CUDA tensors
and on NumPy arrays, the export / import pair, and size-independent properties at a size no CPU oracle is run at."""
import ctypes as C

import numpy as np
import pytest

from oracle import stochqn_np as O
from oracle.driver import HostStepper, run_trace
from oracle.problems import Logistic, Rosenbrock
from stochqn_b200 import _lib
from stochqn_b200.optimizers import SQN_free, adaQN_free, oLBFGS_free

pytestmark = pytest.mark.gpu


def _drive(opt, x, prob, ncalls, step, to_np):
    seq = []
    for _ in range(ncalls):
        req = opt.run_optimizer(x, step)
        seq.append((req["task"], req["info"]["x_changed_in_run"], req["info"]["iteration_number"], req["info"]["iteration_info"]))
        task = req["task"]
        if task in ("calc_grad", "calc_grad_same_batch", "calc_grad_big_batch"):
            kind = {"calc_grad": "new", "calc_grad_same_batch": "same", "calc_grad_big_batch": "big"}[task]
            opt.update_gradient(prob.grad(to_np(req["requested_on"]), kind))
        elif task == "calc_hess_vec":
            opt.update_hess_vec(prob.hess_vec(to_np(req["requested_on"][0]), to_np(req["requested_on"][1])))
        elif task == "calc_fun_val_batch":
            opt.update_function(prob.fun(to_np(req["requested_on"])))
    return seq


@pytest.mark.parametrize("where", ["cuda", "numpy"])
def test_free_mode_classes_follow_the_oracle(where):
    import torch
    names = {101: "calc_grad", 102: "calc_grad_same_batch", 103: "calc_grad_big_batch", 104: "calc_hess_vec", 105: "calc_fun_val_batch"}
    infos = {200: "no_problems_encountered", 201: "func_increased", 202: "curvature_too_small", 203: "search_direction_was_nan"}
    specs = [
        (oLBFGS_free(mem_size=5, min_curvature=1e-4), O.OracleOLBFGS(40, 5, 0.0, 0.0, 1e-4, 1, 1), 1e-1),
        (SQN_free(mem_size=5, bfgs_upd_freq=5, min_curvature=1e-4), O.OracleSQN(40, 5, 5, 1e-4, 0, 0.0, 1, 1), 1e-1),
        (adaQN_free(mem_size=5, fisher_size=20, bfgs_upd_freq=5, max_incr=1.01, min_curvature=1e-4, scal_reg=1e-4, rmsprop_weight=0.9),
         O.OracleAdaQN(40, 5, 20, 5, 1.01, 1e-4, 1e-4, 0.9, 0, 0.0, 1, 1), 1e-2),
    ]
    for opt, oracle, step in specs:
        p1, p2 = Logistic(), Logistic()
        if where == "cuda":
            x = torch.tensor(p1.x0(), device="cuda", dtype=torch.float64)
            to_np = lambda t: t.detach().cpu().numpy()       # noqa: E731
        else:
            x = p1.x0().copy()
            to_np = lambda a: np.array(a, dtype=np.float64)  # noqa: E731
        seq = _drive(opt, x, p1, 120, step, to_np)
        tr = run_trace(HostStepper(oracle, p2.x0()), p2, 120, step, keep_x=True)
        want = [(names[r["task"]], bool(r["ret"]), r["niter"], infos[r["info"]]) for r in tr]
        assert seq == want
        xf = to_np(x)
        assert np.max(np.abs(xf - tr[-1]["x"])) <= 1e-10 * np.max(np.abs(tr[-1]["x"]))


def test_wrong_dtype_and_layout_are_refused_like_the_reference():
    import torch
    opt = oLBFGS_free()
    with pytest.raises(ValueError):
        opt.run_optimizer(np.zeros(4, dtype=np.float32), 1e-3)           # reference: "'x' has wrong dtype."
    with pytest.raises(ValueError):
        opt.run_optimizer(torch.zeros(4, device="cuda", dtype=torch.float32), 1e-3)
    with pytest.raises(AssertionError):
        oLBFGS_free(mem_size=0)
    with pytest.raises(AssertionError):
        adaQN_free(rmsprop_weight=1.5)


def test_export_import_round_trip_resumes_bit_identically():
    """Checkpoint / resume: export the device state in the reference's layout, import it into a
    fresh workspace, continue - the continuation must equal the uninterrupted run bit for bit."""
    import torch
    abi = _lib.load(np.float64)
    lib = abi.lib
    n, m = 1001, 6
    prob = Rosenbrock(n)

    def make():
        ws = lib.initialize_oLBFGS(n, m, 0.0, 0.0, 1e-4, 1, 1)
        x = torch.tensor(prob.x0(), device="cuda", dtype=torch.float64)
        g = torch.zeros_like(x)
        return ws, x, g

    def run(ws, x, g, ncalls):
        req, task, info = C.c_void_p(), C.c_int(), C.c_int()
        for _ in range(ncalls):
            lib.run_oLBFGS(1e-4, x.data_ptr(), g.data_ptr(), C.byref(req), C.byref(task), ws, C.byref(info))
            g.copy_(torch.from_numpy(prob.grad(x.cpu().numpy())))

    wa, xa, ga = make()
    run(wa, xa, ga, 41)                     # odd number of calls: stops between a step and its pair update
    S = np.zeros((m, n)); Y = np.zeros((m, n)); gp = np.zeros(n)
    hs = _lib.HostState(s_mem=S.ctypes.data, y_mem=Y.ctypes.data, grad_prev=gp.ctypes.data)
    assert lib.stochqn_b200_export(wa, C.byref(hs)) == 0
    assert np.any(S != 0) and np.any(Y != 0)
    wb, xb, gb = make()
    xb.copy_(xa); gb.copy_(ga)
    for f in ("niter", "section"):
        setattr(wb.contents, f, getattr(wa.contents, f))
    for f in ("mem_used", "mem_st_ix"):
        setattr(wb.contents.bfgs_memory.contents, f, getattr(wa.contents.bfgs_memory.contents, f))
    assert lib.stochqn_b200_import(wb, C.byref(hs)) == 0
    run(wa, xa, ga, 30)
    run(wb, xb, gb, 30)
    assert wa.contents.niter == wb.contents.niter
    # the resumed Gram matrices are recomputed from the pairs (different summation history): tolerance, not bits
    assert torch.max(torch.abs(xa - xb)).item() <= 1e-12 * torch.max(torch.abs(xa)).item()
    lib.dealloc_oLBFGS(wa); lib.dealloc_oLBFGS(wb)


def test_direction_is_linear_in_the_gradient_at_large_n():
    """Size-independent property at n = 2^22 (no CPU oracle at this size): with fixed pairs H is a linear operator,
    so the step taken for g1 + 2*g2 equals step(g1) + 2*step(g2); and s'y / Gram bookkeeping survives wrap-around."""
    import torch
    abi = _lib.load(np.float64)
    lib = abi.lib
    n, m = 2 ** 22 + 5, 4
    gen = torch.Generator(device="cuda").manual_seed(1)

    def fill(ws, x, g):
        req, task, info = C.c_void_p(), C.c_int(), C.c_int()
        lib.run_oLBFGS(1e-3, x.data_ptr(), g.data_ptr(), C.byref(req), C.byref(task), ws, C.byref(info))
        a = torch.linspace(1.0, 2.0, n, device="cuda", dtype=torch.float64)
        for _ in range(2 * (m + 2)):         # quadratic f = 0.5 a.x^2: gradient a*x, pairs have positive curvature
            g.copy_(a * x)
            lib.run_oLBFGS(1e-3, x.data_ptr(), g.data_ptr(), C.byref(req), C.byref(task), ws, C.byref(info))
        assert ws.contents.bfgs_memory.contents.mem_used == m and task.value == 101
        return req, task, info

    outs = []
    g1 = torch.randn(n, device="cuda", dtype=torch.float64, generator=gen)
    g2 = torch.randn(n, device="cuda", dtype=torch.float64, generator=gen)
    for gv in (g1, g2, g1 + 2 * g2):
        ws = lib.initialize_oLBFGS(n, m, 0.0, 0.0, 0.0, 1, 1)
        x = torch.ones(n, device="cuda", dtype=torch.float64)
        g = torch.zeros_like(x)
        req, task, info = fill(ws, x, g)
        x0 = x.clone()
        g.copy_(gv)
        lib.run_oLBFGS(1.0, x.data_ptr(), g.data_ptr(), C.byref(req), C.byref(task), ws, C.byref(info))
        assert info.value == 200
        outs.append(x0 - x)                  # = H g  (step size 1)
        lib.dealloc_oLBFGS(ws)
    lin = outs[0] + 2 * outs[1]
    assert torch.max(torch.abs(outs[2] - lin)).item() <= 1e-11 * torch.max(torch.abs(lin)).item()
