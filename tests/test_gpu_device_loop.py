"""GPU: the device-side request loop (csrc/kernels_loop.cuh, stochqn_b200_fit_batches) against the host-driven loop of
stochqn_b200_fit_batch on the same mini-batches: the ring counters never visit the host between mini-batches, the
step / pair kernels take the decisions of src/stochqn.c:825-835 (direction accepted?) and 883-900 (curvature) themselves.
Same call tallies (by info code), same counters, same iterate (the arithmetic of the update is K3's, bit for bit; the dot
products are summed in another order: 1e-10) - also when events are frequent (a curvature threshold that rejects
most pairs: quirk Q1 zeroes a slot of the full memory, the next direction is NaN, the memory is flushed)."""
import ctypes as C

import numpy as np
import pytest

from stochqn_b200 import _lib

pytestmark = pytest.mark.gpu


def _problem(torch, nrows, n, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    X = torch.randn(nrows, n, device="cuda", dtype=torch.float64, generator=g)
    X[:, 0] = 1.0
    w = torch.randn(n, device="cuda", dtype=torch.float64, generator=g) * (2.0 / n ** 0.5)
    y = (torch.rand(nrows, device="cuda", dtype=torch.float64, generator=g) < torch.sigmoid(X @ w)).to(torch.float64)
    return X, y


def _run(kind, n, batch, nbatches, min_curv, loop_max_n, chunk, step=0.1, L=4, mem=5):
    import torch
    abi = _lib.load(np.float64)
    lib = abi.lib
    nrows = batch * nbatches
    X, y = _problem(torch, nrows, n, 11)
    x = torch.zeros(n, device="cuda", dtype=torch.float64)
    big = batch * L
    work = torch.empty(lib.stochqn_b200_logistic_work_size(max(batch, big), n), device="cuda", dtype=torch.uint8)
    if kind == "oLBFGS":
        ws = lib.initialize_oLBFGS(n, mem, 0.0, 0.0, min_curv, 1, 1)
    else:
        ws = lib.initialize_SQN(n, mem, L, min_curv, 0, 0.0, 1, 1)
    assert ws
    assert lib.stochqn_b200_set_option(ws, _lib.OPT_DEVICE_LOOP_MAX_N, loop_max_n) == 0
    req, req_vec, task, info = C.c_void_p(), C.c_void_p(), C.c_int(), C.c_int()
    g0 = torch.zeros(n, device="cuda", dtype=torch.float64)
    if kind == "oLBFGS":
        lib.run_oLBFGS(step, x.data_ptr(), g0.data_ptr(), C.byref(req), C.byref(task), ws, C.byref(info))
    else:
        lib.run_SQN(step, x.data_ptr(), g0.data_ptr(), g0.data_ptr(), C.byref(req), C.byref(req_vec), C.byref(task), ws, C.byref(info))
    assert task.value == 101
    M = abi.Model(0, 0, n, 0, 1e-5, work.data_ptr())
    data = _lib.Rows(X.data_ptr(), n, y.data_ptr(), 1, None, nrows)
    tally = dict(calls=0, n_info=[0, 0, 0, 0], x_changed=0)
    rep = _lib.FitReport()
    b = 0
    while b < nbatches:
        cnt = min(chunk, nbatches - b)
        LL = C.c_longlong * cnt
        lf, lr = LL(), LL()
        for i in range(cnt):
            e = (b + i + 1) * batch
            lr[i] = min(big, e)
            lf[i] = e - lr[i]
        rc = lib.stochqn_b200_fit_batches(ws, x.data_ptr(), step, C.byref(M), C.byref(data), b * batch, batch, cnt, lf, lr, None,
                                          C.byref(task), C.byref(req), C.byref(req_vec), C.byref(rep))
        assert rc == 0, _lib.last_error(abi)
        assert task.value == 101 and req.value == x.data_ptr()
        tally["calls"] += rep.calls
        tally["x_changed"] += rep.x_changed
        for i in range(4):
            tally["n_info"][i] += rep.n_info[i]
        b += cnt
    w = ws.contents
    m = w.bfgs_memory.contents
    out = dict(tally, niter=int(w.niter), section=int(w.section), mem_used=int(m.mem_used), mem_st_ix=int(m.mem_st_ix),
               loop_steps=_lib.get_stat(abi, ws, _lib.STAT_DEVICE_LOOP_STEPS), x=x.cpu().numpy())
    {"oLBFGS": lib.dealloc_oLBFGS, "SQN": lib.dealloc_SQN}[kind](ws)
    return out


@pytest.mark.parametrize("kind", ["oLBFGS", "SQN"])
@pytest.mark.parametrize("n", [37, 1001, 2049, 5000])        # one 1024-thread CTA up to 2048, a cooperative grid above
@pytest.mark.parametrize("factor", [0.0, 0.8, 3.0])
def test_device_loop_matches_the_host_driven_loop(kind, n, factor):
    # curvature of a pair on a 64-row batch is about 0.2 * max(1, n / 64) (sigmoid'(z) ~ 0.2, |x_i|^2 ~ n): a threshold at
    # 0.8 of that rejects some pairs, at 3 times that every pair (quirk Q1: the slot is zeroed, with a full memory the
    # next direction is NaN and the memory is flushed)
    min_curv = 1e-4 if factor == 0 else factor * 0.2 * max(1.0, n / 64.0)
    a = _run(kind, n, 64, 45, min_curv, loop_max_n=1 << 16, chunk=7)
    b = _run(kind, n, 64, 45, min_curv, loop_max_n=0, chunk=7)
    assert a["loop_steps"] > 0 and b["loop_steps"] == 0
    for k in ("calls", "n_info", "niter", "section", "mem_used", "mem_st_ix"):
        assert a[k] == b[k], (k, a[k], b[k])
    if factor > 2:
        assert a["n_info"][2] > 0, "the forced curvature rejections did not happen"
    assert np.all(np.isfinite(a["x"]))
    assert np.max(np.abs(a["x"] - b["x"])) <= 1e-10 * max(np.max(np.abs(b["x"])), 1e-300)


def test_chunking_of_the_calls_does_not_matter():
    a = _run("oLBFGS", 1001, 64, 40, 1e-4, loop_max_n=1 << 16, chunk=40)
    b = _run("oLBFGS", 1001, 64, 40, 1e-4, loop_max_n=1 << 16, chunk=1)
    assert np.array_equal(a["x"], b["x"]) and a["n_info"] == b["n_info"] and a["mem_st_ix"] == b["mem_st_ix"]
