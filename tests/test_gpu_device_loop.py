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


def _run(kind, n, batch, nbatches, min_curv, loop_max_n, chunk, step=0.1, L=4, mem=5, fused=1, model=0, weights=False, short_tail=0,
         dtype=np.float64):
    """model 0: R conventions (the intercept is column 0 of X); model 1: scikit-learn conventions with an unpenalised
    intercept stored last (n - 1 columns, labels +-1).  short_tail: rows missing from the last mini-batch."""
    import torch
    abi = _lib.load(dtype)
    lib = abi.lib
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    nrows = batch * nbatches - short_tail
    X, y = _problem(torch, nrows, n, 11)
    sw = None
    if weights:
        sw = (0.5 + torch.rand(nrows, device="cuda", dtype=torch.float64, generator=torch.Generator(device="cuda").manual_seed(5))).to(tdt)
    ncols = n
    if model == 1:
        X = X[:, 1:].contiguous()
        y = 2.0 * y - 1.0
        ncols = n - 1
    X, y = X.to(tdt), y.to(tdt)
    x = torch.zeros(n, device="cuda", dtype=tdt)
    big = batch * L
    work = torch.empty(lib.stochqn_b200_logistic_work_size(max(batch, big), n), device="cuda", dtype=torch.uint8)
    if kind == "oLBFGS":
        ws = lib.initialize_oLBFGS(n, mem, 0.0, 0.0, min_curv, 1, 1)
    elif kind == "SQN":
        ws = lib.initialize_SQN(n, mem, L, min_curv, 0, 0.0, 1, 1)
    else:       # adaQN_fisher: AdaGrad + empirical Fisher pairs; adaQN_gd: RMSProp + gradient differences; adaQN_incr: Fisher + max_incr
        fisher = 0 if kind == "adaQN_gd" else 7
        ws = lib.initialize_adaQN(n, mem, max(fisher, 1) if fisher else 1, L, 1.01 if kind == "adaQN_incr" else 0.0, min_curv, 1e-4,
                                  0.9 if kind == "adaQN_gd" else 0.0, 1 if kind == "adaQN_gd" else 0, 0.0, 1, 1)
    assert ws
    assert lib.stochqn_b200_set_option(ws, _lib.OPT_DEVICE_LOOP_MAX_N, loop_max_n) == 0
    assert lib.stochqn_b200_set_option(ws, _lib.OPT_FUSED_FIT, fused) == 0
    req, req_vec, task, info = C.c_void_p(), C.c_void_p(), C.c_int(), C.c_int()
    g0 = torch.zeros(n, device="cuda", dtype=tdt)
    if kind == "oLBFGS":
        lib.run_oLBFGS(step, x.data_ptr(), g0.data_ptr(), C.byref(req), C.byref(task), ws, C.byref(info))
    elif kind == "SQN":
        lib.run_SQN(step, x.data_ptr(), g0.data_ptr(), g0.data_ptr(), C.byref(req), C.byref(req_vec), C.byref(task), ws, C.byref(info))
    else:
        lib.run_adaQN(step, x.data_ptr(), 0.0, g0.data_ptr(), C.byref(req), C.byref(task), ws, C.byref(info))
    assert task.value == 101
    M = abi.Model(model, 1 if model == 1 else 0, ncols, 0, 1e-5 if model == 0 else 1e-2, work.data_ptr())
    data = _lib.Rows(X.data_ptr(), ncols, y.data_ptr(), 1, sw.data_ptr() if sw is not None else None, nrows)
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
    fisher = (0, 0)
    if kind.startswith("adaQN") and w.fisher_memory:
        fisher = (int(w.fisher_memory.contents.mem_used), int(w.fisher_memory.contents.mem_st_ix))
    out = dict(tally, niter=int(w.niter), section=int(w.section), mem_used=int(m.mem_used), mem_st_ix=int(m.mem_st_ix), fisher=fisher,
               loop_steps=_lib.get_stat(abi, ws, _lib.STAT_DEVICE_LOOP_STEPS), fit_steps=_lib.get_stat(abi, ws, _lib.STAT_FUSED_FIT_STEPS),
               x=x.double().cpu().numpy())
    {"oLBFGS": lib.dealloc_oLBFGS, "SQN": lib.dealloc_SQN}.get(kind, lib.dealloc_adaQN)(ws)
    return out


@pytest.mark.parametrize("fused", [0, 1])                    # 0: one kernel sequence per mini-batch (kernels_loop.cuh); 1: one launch per run (kernels_fit.cuh)
@pytest.mark.parametrize("kind", ["oLBFGS", "SQN"])
@pytest.mark.parametrize("n", [37, 1001, 2049, 5000])        # one 1024-thread CTA up to 2048, a cooperative grid above
@pytest.mark.parametrize("factor", [0.0, 0.8, 3.0])
def test_device_loop_matches_the_host_driven_loop(kind, n, factor, fused):
    # curvature of a pair on a 64-row batch is about 0.2 * max(1, n / 64) (sigmoid'(z) ~ 0.2, |x_i|^2 ~ n): a threshold at
    # 0.8 of that rejects some pairs, at 3 times that every pair (quirk Q1: the slot is zeroed, with a full memory the
    # next direction is NaN and the memory is flushed)
    min_curv = 1e-4 if factor == 0 else factor * 0.2 * max(1.0, n / 64.0)
    a = _run(kind, n, 64, 45, min_curv, loop_max_n=1 << 16, chunk=7, fused=fused)
    b = _run(kind, n, 64, 45, min_curv, loop_max_n=0, chunk=7, fused=fused)
    assert a["loop_steps"] > 0 and b["loop_steps"] == 0
    assert (a["fit_steps"] > 0) == bool(fused) and b["fit_steps"] == 0
    for k in ("calls", "n_info", "niter", "section", "mem_used", "mem_st_ix"):
        assert a[k] == b[k], (k, a[k], b[k])
    if factor > 2:
        assert a["n_info"][2] > 0, "the forced curvature rejections did not happen"
    assert np.all(np.isfinite(a["x"]))
    assert np.max(np.abs(a["x"] - b["x"])) <= 1e-10 * max(np.max(np.abs(b["x"])), 1e-300)


@pytest.mark.parametrize("fused", [0, 1])
def test_chunking_of_the_calls_does_not_matter(fused):
    a = _run("oLBFGS", 1001, 64, 40, 1e-4, loop_max_n=1 << 16, chunk=40, fused=fused)
    b = _run("oLBFGS", 1001, 64, 40, 1e-4, loop_max_n=1 << 16, chunk=1, fused=fused)
    assert np.array_equal(a["x"], b["x"]) and a["n_info"] == b["n_info"] and a["mem_st_ix"] == b["mem_st_ix"]


@pytest.mark.parametrize("kind", ["oLBFGS", "SQN"])
@pytest.mark.parametrize("model,weights,short_tail,batch,n", [
    (0, True, 0, 1000, 1001),      # BASELINE config-1 shape with sample weights
    (1, False, 0, 300, 777),       # scikit-learn conventions: labels +-1, sums, unpenalised intercept stored last
    (1, True, 17, 129, 4097),      # widest register layout, ragged last mini-batch
    (0, False, 5, 64, 5120),       # the largest n the fused kernel takes
    (0, False, 0, 7, 3),           # fewer rows and variables than CTAs
])
def test_fused_runs_on_other_models_and_shapes(kind, model, weights, short_tail, batch, n):
    kw = dict(model=model, weights=weights, short_tail=short_tail)
    a = _run(kind, n, batch, 23, 1e-4, loop_max_n=1 << 16, chunk=9, fused=1, **kw)
    b = _run(kind, n, batch, 23, 1e-4, loop_max_n=0, chunk=9, fused=0, **kw)
    assert a["fit_steps"] == 23 - (23 // 4 if kind == "SQN" else 0) and b["loop_steps"] == 0
    for k in ("calls", "n_info", "niter", "section", "mem_used", "mem_st_ix"):
        assert a[k] == b[k], (k, a[k], b[k])
    assert np.all(np.isfinite(a["x"]))
    assert np.max(np.abs(a["x"] - b["x"])) <= 1e-10 * max(np.max(np.abs(b["x"])), 1e-300)


def test_fused_run_in_single_precision():
    a = _run("oLBFGS", 1001, 500, 20, 1e-4, loop_max_n=1 << 16, chunk=20, fused=1, dtype=np.float32)
    b = _run("oLBFGS", 1001, 500, 20, 1e-4, loop_max_n=0, chunk=20, fused=0, dtype=np.float32)
    assert a["fit_steps"] == 20
    for k in ("calls", "n_info", "niter", "mem_used", "mem_st_ix"):
        assert a[k] == b[k], (k, a[k], b[k])
    assert np.max(np.abs(a["x"] - b["x"])) <= 1e-4 * np.max(np.abs(b["x"]))


@pytest.mark.parametrize("kind", ["adaQN_fisher", "adaQN_gd", "adaQN_incr"])
@pytest.mark.parametrize("n,batch,nbatches,mem", [(37, 64, 45, 5), (1001, 64, 45, 5), (5000, 32, 45, 12), (30011, 16, 26, 14)])
@pytest.mark.parametrize("factor", [0.0, 0.5, 3.0])
def test_adaqn_one_launch_steps_match_the_host_driven_loop(kind, n, batch, nbatches, mem, factor):
    """The ordinary steps of adaQN as ONE launch each (kl_ada, csrc/kernels_loop.cuh: stochqn.c:802-840 with 720-783) inside
    stochqn_b200_fit_batches, against the five-launch host-driven route: same tallies, ring and Fisher counters, iterate.
    factor 3: a curvature threshold that rejects every pair (Fisher pairs have curvature ~ |g's|^2/|s|^2, grad-diff pairs
    as in the oLBFGS test above)."""
    min_curv = 1e-4 if factor == 0 else factor * 0.2 * max(1.0, n / float(batch))
    step = 0.05
    a = _run(kind, n, batch, nbatches, min_curv, loop_max_n=1 << 19, chunk=7, step=step, mem=mem)
    b = _run(kind, n, batch, nbatches, min_curv, loop_max_n=0, chunk=7, step=step, mem=mem)
    assert a["loop_steps"] > 0 and b["loop_steps"] == 0
    for k in ("calls", "n_info", "niter", "section", "mem_used", "mem_st_ix", "fisher"):
        assert a[k] == b[k], (k, a[k], b[k])
    assert np.all(np.isfinite(a["x"]))
    assert np.max(np.abs(a["x"] - b["x"])) <= 1e-10 * max(np.max(np.abs(b["x"])), 1e-300)


def _run_multinomial(kind, d, K, batch, nbatches, loop_max_n, fused, chunk, mem=5, L=4, short_tail=0, weights=True, min_curv=1e-4):
    """adaQN on the multinomial model (model 2: one-hot labels, intercept last) through stochqn_b200_fit_batches."""
    import torch
    abi = _lib.load(np.float64)
    lib = abi.lib
    n = K * (d + 1)
    nrows = batch * nbatches - short_tail
    g = torch.Generator(device="cuda").manual_seed(21)
    X = torch.randn(nrows, d, device="cuda", dtype=torch.float64, generator=g) / d ** 0.5
    lab = torch.randint(0, K, (nrows,), device="cuda", generator=g)
    Y = torch.zeros(nrows, K, device="cuda", dtype=torch.float64)
    Y[torch.arange(nrows, device="cuda"), lab] = 1.0
    sw = (0.5 + torch.rand(nrows, device="cuda", dtype=torch.float64, generator=g)) / batch if weights else None
    x = torch.randn(n, device="cuda", dtype=torch.float64, generator=g) * 0.1
    big = batch * L
    work = torch.empty(lib.stochqn_b200_multinomial_work_size(max(batch, big), d, K), device="cuda", dtype=torch.uint8)
    fisher = 0 if kind == "adaQN_gd" else 7
    ws = lib.initialize_adaQN(n, mem, fisher if fisher else 1, L, 0.0, min_curv, 1e-4, 0.9 if kind == "adaQN_gd" else 0.0, 1 if kind == "adaQN_gd" else 0,
                              0.0, 1, 1)
    assert ws
    assert lib.stochqn_b200_set_option(ws, _lib.OPT_DEVICE_LOOP_MAX_N, loop_max_n) == 0
    assert lib.stochqn_b200_set_option(ws, _lib.OPT_FUSED_FIT, fused) == 0
    req, req_vec, task, info = C.c_void_p(), C.c_void_p(), C.c_int(), C.c_int()
    g0 = torch.zeros(n, device="cuda", dtype=torch.float64)
    lib.run_adaQN(0.05, x.data_ptr(), 0.0, g0.data_ptr(), C.byref(req), C.byref(task), ws, C.byref(info))
    M = abi.Model(2, 1, d, K, 1e-2, work.data_ptr())
    data = _lib.Rows(X.data_ptr(), d, Y.data_ptr(), K, sw.data_ptr() if sw is not None else None, nrows)
    rep = _lib.FitReport()
    tally = dict(calls=0, n_info=[0, 0, 0, 0])
    launches0 = _lib.launch_count()
    b = 0
    while b < nbatches:
        cnt = min(chunk, nbatches - b)
        LL = C.c_longlong * cnt
        lf, lr = LL(), LL()
        for i in range(cnt):
            e = min((b + i + 1) * batch, nrows)
            lr[i] = min(big, e)
            lf[i] = e - lr[i]
        rc = lib.stochqn_b200_fit_batches(ws, x.data_ptr(), 0.05, C.byref(M), C.byref(data), b * batch, batch, cnt, lf, lr, None,
                                          C.byref(task), C.byref(req), C.byref(req_vec), C.byref(rep))
        assert rc == 0, _lib.last_error(abi)
        tally["calls"] += rep.calls
        for i in range(4):
            tally["n_info"][i] += rep.n_info[i]
        b += cnt
    torch.cuda.synchronize()
    w = ws.contents
    m = w.bfgs_memory.contents
    fm = (int(w.fisher_memory.contents.mem_used), int(w.fisher_memory.contents.mem_st_ix)) if w.fisher_memory else (0, 0)
    out = dict(tally, niter=int(w.niter), section=int(w.section), mem_used=int(m.mem_used), mem_st_ix=int(m.mem_st_ix), fisher=fm,
               loop_steps=_lib.get_stat(abi, ws, _lib.STAT_DEVICE_LOOP_STEPS), fit_steps=_lib.get_stat(abi, ws, _lib.STAT_FUSED_FIT_STEPS),
               launches=_lib.launch_count() - launches0, x=x.cpu().numpy())
    lib.dealloc_adaQN(ws)
    return out


@pytest.mark.parametrize("kind", ["adaQN_fisher", "adaQN_gd"])
@pytest.mark.parametrize("d,K,batch,short_tail", [(600, 130, 40, 0), (1836, 159, 50, 13), (333, 250, 17, 0)])
def test_adaqn_multinomial_runs_in_one_launch(kind, d, K, batch, short_tail):
    """adaQN + multinomial model: a run of ordinary mini-batches is ONE persistent launch (kl_fit_mn_ada = mn_grad_small + kl_ada
    per mini-batch, csrc/kernels_loop.cuh), against two launches per mini-batch (FUSED_FIT off) and against the host-driven loop."""
    a = _run_multinomial(kind, d, K, batch, 26, loop_max_n=1 << 19, fused=1, chunk=9, short_tail=short_tail)
    b = _run_multinomial(kind, d, K, batch, 26, loop_max_n=1 << 19, fused=0, chunk=9, short_tail=short_tail)
    c = _run_multinomial(kind, d, K, batch, 26, loop_max_n=0, fused=0, chunk=9, short_tail=short_tail)
    # (the STEP of a boundary mini-batch is taken on the device too; what follows it - averaging, pair - is host work)
    assert a["fit_steps"] == 26 and b["fit_steps"] == 0 and b["loop_steps"] == 26 and c["loop_steps"] == 0
    assert a["launches"] < b["launches"] < c["launches"]
    for k in ("calls", "n_info", "niter", "section", "mem_used", "mem_st_ix", "fisher"):
        assert a[k] == c[k] and b[k] == c[k], (k, a[k], b[k], c[k])
    assert a["mem_used"] > 0 and np.all(np.isfinite(a["x"]))
    scale = max(np.max(np.abs(c["x"])), 1e-300)
    assert np.max(np.abs(a["x"] - c["x"])) <= 1e-10 * scale
    assert np.max(np.abs(b["x"] - c["x"])) <= 1e-10 * scale


def test_adaqn_multinomial_run_with_rejected_pairs():
    """The same with a curvature threshold nobody passes after the memory filled once: quirk Q1 zeroes slots of the full memory, the next
    direction is not finite, the kernel rejects it on the device (x_sum += x, flush) and carries on - tallies and iterate as the host-driven loop."""
    kw = dict(d=600, K=130, batch=40, nbatches=40, chunk=11, mem=3, L=2)
    a = _run_multinomial("adaQN_gd", loop_max_n=1 << 19, fused=1, min_curv=1e6, **kw)
    c = _run_multinomial("adaQN_gd", loop_max_n=0, fused=0, min_curv=1e6, **kw)
    assert a["fit_steps"] == 40 and c["loop_steps"] == 0
    assert c["n_info"][2] > 0, "the forced curvature rejections did not happen"
    for k in ("calls", "n_info", "niter", "section", "mem_used", "mem_st_ix"):
        assert a[k] == c[k], (k, a[k], c[k])
    assert np.all(np.isfinite(a["x"]))
    assert np.max(np.abs(a["x"] - c["x"])) <= 1e-10 * max(np.max(np.abs(c["x"])), 1e-300)
