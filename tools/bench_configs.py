"""Device-resident request loops for the other BASELINE.json configurations (the headline one is bench.py):

  cfg1  oLBFGS  + binary logistic      X 100k x (1000 + intercept), batch 1000, mem_size 10, fp64
  cfg2  SQN (calc_hess_vec) + logistic X 1M x (4096 + intercept), batch 2000, bfgs_upd_freq 10, big batch 20k, fp64
  cfg3  adaQN AdaGrad + Fisher(100)    multinomial 1836 features x 159 classes, batch 50, L 20, max_incr 1.01, fp64
  cfg5  adaQN RMSProp(0.9) + grad-diff multinomial 8192 features x 4096 classes, batch 1024 per GPU, fp32 (tensor cores)

Everything stays on the GPU: the optimizer's requests are served by the bundled device callbacks at `*req`.
Synthetic data (no datasets in the image): X ~ N(0,1)/sqrt(d), labels drawn from a random ground-truth model.
Prints one JSON line per configuration: optimizer steps/s (CUDA events), task / info histograms, final loss.

    python tools/bench_configs.py [cfg1 cfg2 cfg3 cfg5] [--steps N] [--rows-cfg2 R]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stochqn_b200 import _lib


def _gen_rows(nrows, d, dtype, intercept_first, seed, unit_variance=False):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    cols = d + (1 if intercept_first else 0)
    X = torch.empty(nrows, cols, device="cuda", dtype=dtype)
    chunk = max(1, (1 << 28) // cols)
    scale = 1.0 if unit_variance else 1.0 / d ** 0.5
    for r0 in range(0, nrows, chunk):
        r1 = min(nrows, r0 + chunk)
        X[r0:r1, (1 if intercept_first else 0):] = torch.randn(r1 - r0, d, device="cuda", dtype=dtype, generator=g) * scale
    if intercept_first:
        X[:, 0] = 1.0                       # R prepends the intercept column (R/logistic.R:424)
    return X


def run_logistic(name, kind, nrows, d, batch, steps, L=10, big=20000, native=False, quiet=False, cpu_fn=None, step=1e-1):
    """Data as SURVEY.md section 8(d) states it: X ~ N(0,1) (unit variance, intercept column prepended), y ~ Bernoulli(sigma(X w*));
    w* ~ 2 N(0,1)/sqrt(d) keeps the logits O(1), so the labels are noisy, the Hessian is well conditioned (eigenvalues ~ 0.2) and
    the correction pairs pass the curvature test (checked with the reference library on the CPU: no info events in 400 steps)."""
    dtype, tdt, esz = np.float64, torch.float64, 8
    abi = _lib.load(dtype)
    lib = abi.lib
    n = d + 1
    X = _gen_rows(nrows, d, tdt, True, 1, unit_variance=True)
    wtrue = torch.randn(n, device="cuda", dtype=tdt, generator=torch.Generator(device="cuda").manual_seed(2)) * (2.0 / d ** 0.5)
    y = (torch.rand(nrows, device="cuda", dtype=tdt) < torch.sigmoid(X @ wtrue)).to(tdt)
    x = torch.zeros(n, device="cuda", dtype=tdt)
    g = torch.zeros(n, device="cuda", dtype=tdt)
    hv = torch.zeros(n, device="cuda", dtype=tdt)
    loss = torch.zeros(1, device="cuda", dtype=torch.float64)
    work = torch.empty(lib.stochqn_b200_logistic_work_size(max(batch, big), n), device="cuda", dtype=torch.uint8)
    lam = 1e-5                                  # the callbacks' default (R/logistic.R:1,12,23)
    if kind == "oLBFGS":
        ws = lib.initialize_oLBFGS(n, 10, 0.0, 0.0, 1e-4, 1, 1)
    else:
        ws = lib.initialize_SQN(n, 10, L, 1e-4, 0, 0.0, 1, 1)
    assert ws, _lib.last_error(abi)
    req, req_vec, task, info = C.c_void_p(), C.c_void_p(), C.c_int(), C.c_int()
    tasks, infos = {}, {}
    nb = nrows // batch
    state = dict(b=0)

    def call():
        if kind == "oLBFGS":
            lib.run_oLBFGS(step, x.data_ptr(), g.data_ptr(), C.byref(req), C.byref(task), ws, C.byref(info))
        else:
            lib.run_SQN(step, x.data_ptr(), g.data_ptr(), hv.data_ptr(), C.byref(req), C.byref(req_vec), C.byref(task), ws, C.byref(info))
        tasks[task.value] = tasks.get(task.value, 0) + 1
        infos[info.value] = infos.get(info.value, 0) + 1

    def rows(b0, cnt):
        return X.data_ptr() + b0 * n * esz, y.data_ptr() + b0 * esz

    def serve():
        t = task.value
        b = state["b"]
        if t == 101:                              # calc_grad: next batch
            state["b"] = b = (b + 1) % nb
            xp, yp = rows(b * batch, batch)
            lib.stochqn_b200_logistic_grad(xp, n, yp, None, batch, n, req.value, lam, g.data_ptr(), work.data_ptr(), None)
        elif t == 102:                            # calc_grad_same_batch
            xp, yp = rows(b * batch, batch)
            lib.stochqn_b200_logistic_grad(xp, n, yp, None, batch, n, req.value, lam, g.data_ptr(), work.data_ptr(), None)
        elif t == 104:                            # calc_hess_vec on the big batch = the rows of the last L batches
            cnt = min(big, (b + 1) * batch)
            r0 = (b + 1) * batch - cnt
            xp, yp = rows(r0, cnt)
            lib.stochqn_b200_logistic_hess_vec(xp, n, yp, None, cnt, n, req.value, req_vec.value, lam, hv.data_ptr(), work.data_ptr(), None)
        else:
            raise RuntimeError("unexpected task %d" % t)

    call()
    niter = lambda: int(ws.contents.niter)
    warm = 15 * (L if kind == "SQN" else 1)
    if native:
        # the request loop of a mini-batch inside the library: ONE call per mini-batch (stochqn_b200_fit_batch);
        # the long batch (last L mini-batches) is a row range of the resident matrix
        M = abi.Model(0, 0, n, 0, lam, work.data_ptr())
        rep = _lib.FitReport()

        def rows_of(r0, cnt):
            xp, yp = rows(r0, cnt)
            return _lib.Rows(xp, n, yp, 1, None, cnt)

        def serve_call():
            state["b"] = b = (state["b"] + 1) % nb
            rb = rows_of(b * batch, batch)
            cnt = min(big, (b + 1) * batch)
            rl = rows_of((b + 1) * batch - cnt, cnt)
            rc = lib.stochqn_b200_fit_batch(ws, x.data_ptr(), step, C.byref(M), C.byref(rb), C.byref(rl), None, C.byref(task),
                                            C.byref(req), C.byref(req_vec), C.byref(rep))
            assert rc == 0, (rc, _lib.last_error(abi))
            for i in range(4):
                infos[200 + i] = infos.get(200 + i, 0) + rep.n_info[i]
    else:
        def serve_call():
            serve(); call()
    if native == "batches":
        # the request loop on the DEVICE: `chunk` consecutive mini-batches per library call (stochqn_b200_fit_batches), the
        # ring counters come back once per call
        chunk = 50
        data = _lib.Rows(X.data_ptr(), n, y.data_ptr(), 1, None, nrows)
        LL = C.c_longlong * chunk
        lf, lr = LL(), LL()

        def serve_call():
            b0 = (state["b"] + 1) % nb
            cnt_b = min(chunk, nb - b0)
            for i in range(cnt_b):
                e = (b0 + i + 1) * batch
                lr[i] = min(big, e)
                lf[i] = e - lr[i]
            rc = lib.stochqn_b200_fit_batches(ws, x.data_ptr(), step, C.byref(M), C.byref(data), b0 * batch, batch, cnt_b, lf, lr, None,
                                              C.byref(task), C.byref(req), C.byref(req_vec), C.byref(rep))
            assert rc == 0, (rc, _lib.last_error(abi))
            state["b"] = b0 + cnt_b - 1
            for i in range(4):
                infos[200 + i] = infos.get(200 + i, 0) + rep.n_info[i]
    while niter() < warm:
        serve_call()
    torch.cuda.synchronize()
    tasks.clear(); infos.clear()
    launches0 = _lib.launch_count()
    it0 = niter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while niter() < it0 + steps:
        serve_call()
    e1.record()
    torch.cuda.synchronize()
    steps = niter() - it0                     # (a chunked call may run past the requested count)
    ms = e0.elapsed_time(e1) / steps
    xp, yp = rows(0, min(nrows, 20000))
    lib.stochqn_b200_logistic_loss(xp, n, yp, None, min(nrows, 20000), n, x.data_ptr(), lam, loss.data_ptr(), work.data_ptr(), None)
    loop = {"batches": "device (stochqn_b200_fit_batches, 50 mini-batches per call)", True: "library (stochqn_b200_fit_batch)"}.get(
        native, "python (one call per request)")
    out = dict(config=name, optimizer=kind, loop=loop,
               dtype="f64", n=n, rows=nrows, batch=batch, steps=steps, ms_per_step=ms, steps_per_s=1e3 / ms,
               tasks=tasks, infos=infos, mem_used=int(ws.contents.bfgs_memory.contents.mem_used),
               device_loop_steps=_lib.get_stat(abi, ws, _lib.STAT_DEVICE_LOOP_STEPS),
               launches_per_step=(_lib.launch_count() - launches0) / steps, loss_after=float(loss.item()), loss_at_zero=float(np.log(2.0)))
    {"oLBFGS": lib.dealloc_oLBFGS, "SQN": lib.dealloc_SQN}[kind](ws)
    if cpu_fn is not None:          # the reference on the host cores, on rows of the same matrix (bench.py's secondary block)
        out["cpu_reference"] = cpu_fn(X, y)
    if not quiet:
        print(json.dumps(out), flush=True)
    return out


def run_multinomial(name, dtype, d, K, batch, nrows, steps, L, fisher, use_grad_diff, max_incr, rms, step, quiet=False, cpu_fn=None,
                    profile=None, fixed_big=False, native=False):
    """profile "bibtex": the shape AND the set-up of the reference's notebook (example/example_stochqn.ipynb: BibTeX, 1836 binary
    bag-of-words features at 3.75 % density, 159 classes): weights of one per sample (summed loss), reg_param 0.1, start point
    ~ N(0,1) - with these the correction-pair memory fills and stays full (checked with the reference library on the CPU:
    no func_increased / curvature events in 700 steps at step 1e-2).  fixed_big: serve calc_grad_big_batch on the same rows
    every time (a caller's choice; the gradient difference then carries no sampling noise and every pair is accepted)."""
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    esz = 8 if dtype == np.float64 else 4
    abi = _lib.load(dtype)
    lib = abi.lib
    n = K * (d + 1)
    gen = torch.Generator(device="cuda").manual_seed(4)
    if profile == "bibtex":
        X = (torch.rand(nrows, d, device="cuda", dtype=tdt, generator=torch.Generator(device="cuda").manual_seed(3)) < 0.0375).to(tdt)
        lab_scale = 1.0
    else:
        X = _gen_rows(nrows, d, tdt, False, 3)
        lab_scale = 4.0
    Wt = torch.randn(K, d, device="cuda", dtype=tdt, generator=gen)
    lab = torch.empty(nrows, device="cuda", dtype=torch.int32)
    for r0 in range(0, nrows, 4096):
        r1 = min(nrows, r0 + 4096)
        lab[r0:r1] = torch.argmax(X[r0:r1] @ Wt.T * lab_scale + torch.randn(r1 - r0, K, device="cuda", dtype=tdt, generator=gen), dim=1).to(torch.int32)
    del Wt
    nval = min(nrows, 740)
    big = min(nrows, batch * L)
    x = torch.zeros(n, device="cuda", dtype=tdt)
    if profile == "bibtex":
        x = torch.randn(n, device="cuda", dtype=tdt, generator=gen)
    x0_host = x.cpu().numpy().copy()
    g = torch.zeros(n, device="cuda", dtype=tdt)
    loss = torch.zeros(1, device="cuda", dtype=torch.float64)
    work = torch.empty(lib.stochqn_b200_multinomial_work_size(max(batch, big, nval), d, K), device="cuda", dtype=torch.uint8)
    if profile == "bibtex":
        sw = {c: torch.ones(c, device="cuda", dtype=tdt) for c in {batch, big, nval}}              # summed log-loss (notebook)
        alpha = 1e-1
    else:
        sw = {c: torch.full((c,), 1.0 / c, device="cuda", dtype=tdt) for c in {batch, big, nval}}     # mean log-loss
        alpha = 1e-3
    ws = lib.initialize_adaQN(n, 10, max(fisher, 1), L, max_incr, 1e-4, 1e-4, rms, use_grad_diff, 0.0, 1, 1)
    assert ws, _lib.last_error(abi)
    req, task, info = C.c_void_p(), C.c_int(), C.c_int()
    tasks, infos = {}, {}
    nb = nrows // batch
    state = dict(b=0, f=0.0)

    def call():
        lib.run_adaQN(step, x.data_ptr(), state["f"], g.data_ptr(), C.byref(req), C.byref(task), ws, C.byref(info))
        tasks[task.value] = tasks.get(task.value, 0) + 1
        infos[info.value] = infos.get(info.value, 0) + 1

    def grad_on(r0, cnt, want_loss=False):
        return lib.stochqn_b200_multinomial_loss_grad(X.data_ptr() + r0 * d * esz, d, None, K, lab.data_ptr() + r0 * 4, sw[cnt].data_ptr(), cnt, d, K, 1,
                                                      req.value, alpha, None if want_loss else g.data_ptr(), loss.data_ptr() if want_loss else None,
                                                      work.data_ptr(), None)

    def serve():
        t = task.value
        b = state["b"]
        if t == 101:
            state["b"] = b = (b + 1) % nb
            assert grad_on(b * batch, batch) == 0
        elif t == 103:                            # calc_grad_big_batch: the rows of the last L batches (or always the same rows)
            cnt = big
            r0 = 0 if fixed_big else max(0, (b + 1) * batch - cnt)
            assert grad_on(r0, cnt) == 0
        elif t == 105:                            # calc_fun_val_batch: validation rows
            assert grad_on(0, nval, want_loss=True) == 0
            state["f"] = float(loss.item())
        else:
            raise RuntimeError("unexpected task %d" % t)

    call()
    niter = lambda: int(ws.contents.niter)
    warm = 12 * L                 # mem_size + 2 correction pairs: the timed steps run with the memory full

    def serve_call():
        serve(); call()
    if native:
        # the request loop inside the library (stochqn_b200_fit_batches): `chunk` consecutive mini-batches per call; the ordinary
        # steps are two launches each (mn_grad_small + kl_ada) with no host wait, the pair iterations take the host-driven loop
        assert profile == "bibtex" and not fixed_big, "one weight array must serve every batch size"
        chunk = 50
        Yoh = torch.zeros(nrows, K, device="cuda", dtype=tdt)
        Yoh[torch.arange(nrows, device="cuda"), lab.long()] = 1.0
        swall = torch.ones(nrows, device="cuda", dtype=tdt)
        M = abi.Model(2, 1, d, K, alpha, work.data_ptr())
        data = _lib.Rows(X.data_ptr(), d, Yoh.data_ptr(), K, swall.data_ptr(), nrows)
        val = _lib.Rows(X.data_ptr(), d, Yoh.data_ptr(), K, swall.data_ptr(), nval)
        rep = _lib.FitReport()
        LL = C.c_longlong * chunk
        lf, lr = LL(), LL()
        req_vec = C.c_void_p()

        def serve_call():
            b0 = (state["b"] + 1) % nb
            cnt_b = min(chunk, nb - b0)
            for i in range(cnt_b):
                e = (b0 + i + 1) * batch
                lr[i] = min(big, e)
                lf[i] = e - lr[i]
            rc = lib.stochqn_b200_fit_batches(ws, x.data_ptr(), step, C.byref(M), C.byref(data), b0 * batch, batch, cnt_b, lf, lr, C.byref(val),
                                              C.byref(task), C.byref(req), C.byref(req_vec), C.byref(rep))
            assert rc == 0, (rc, _lib.last_error(abi))
            state["b"] = b0 + cnt_b - 1
            for i in range(4):
                infos[200 + i] = infos.get(200 + i, 0) + rep.n_info[i]
    while niter() < warm:
        serve_call()
    torch.cuda.synchronize()
    tasks.clear(); infos.clear()
    launches0 = _lib.launch_count()
    it0 = niter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while niter() < it0 + steps:
        serve_call()
    e1.record()
    torch.cuda.synchronize()
    steps = niter() - it0
    ms = e0.elapsed_time(e1) / steps
    req.value = x.data_ptr()
    grad_on(0, nval, want_loss=True)
    out = dict(config=name, optimizer="adaQN", loop="device (stochqn_b200_fit_batches)" if native else "python (one call per request)",
               device_loop_steps=_lib.get_stat(abi, ws, _lib.STAT_DEVICE_LOOP_STEPS), dtype="f64" if esz == 8 else "f32", n=n, features=d, classes=K, batch=batch, steps=steps,
               ms_per_step=ms, steps_per_s=1e3 / ms, tasks=tasks, infos=infos, mem_used=int(ws.contents.bfgs_memory.contents.mem_used),
               launches_per_step=(_lib.launch_count() - launches0) / steps, loss_after=float(loss.item()), loss_at_zero=float(np.log(K)))
    lib.dealloc_adaQN(ws)
    if cpu_fn is not None:
        out["cpu_reference"] = cpu_fn(X, lab, x0_host)
    if not quiet:
        print(json.dumps(out), flush=True)
    return out


def run_multinomial_sharded(name, dtype, d, K, batch_per_gpu, nrows_per_gpu, steps, L, rms, step, mode="zero1", warm_cycles=3,
                            return_x=False, quiet=False, union_world=None, fixed_big=False):
    """BASELINE config 5 over several GPUs (launch with torch.distributed.run, one rank per GPU): batch ROWS shard across
    the ranks, each rank evaluates the multinomial gradient on its rows (weights 1/global batch), then either
      mode "allreduce": ncclAllReduce of the n-vector, every rank runs the same (replicated) adaQN step, or
      mode "zero1"    : ncclReduceScatter -> rank r steps block r of the optimizer state (the library's sharded optimizer:
                        dot partials exchanged inside the solve kernels) -> ncclAllGather of the point the next request names.
    adaQN with RMSProp + gradient differencing (calc_grad / calc_grad_big_batch), max_incr 0."""
    import torch.distributed as dist
    from stochqn_b200.distributed import init_comm

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if union_world:
        # the UNSHARDED twin of a run over `union_world` ranks, on this GPU alone: global batch b = the rows every rank
        # holds for its batch b, one after the other (called by one rank only; no collective inside)
        rank, world = 0, 1
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    esz = 8 if dtype == np.float64 else 4
    abi = _lib.load(dtype)
    lib = abi.lib
    comm = init_comm(abi, rank, world) if world > 1 else None
    if comm is not None and mode in ("fused", "p2p") and not lib.stochqn_b200_comm_uses_p2p(comm):      # the same answer on every rank
        lib.stochqn_b200_comm_destroy(comm)
        raise RuntimeError("mode %s needs the peer-memory path (no peer access between the devices, or STOCHQN_B200_NO_P2P is set)" % mode)
    n = K * (d + 1)
    assert n % world == 0, "n must divide by the number of ranks for reduce-scatter"
    sharded_opt = mode in ("zero1", "fused", "p2p")
    blk = n // world if sharded_opt else n
    off = rank * blk if sharded_opt else 0
    # this rank's rows of every global batch: global row = b * (world * batch_per_gpu) + rank * batch_per_gpu + i
    nb = nrows_per_gpu // batch_per_gpu
    gen = torch.Generator(device="cuda").manual_seed(100)
    Wt = torch.randn(K, d, device="cuda", dtype=tdt, generator=gen)                    # same ground truth on every rank
    if union_world:
        W_ = int(union_world)
        parts = [torch.randn(nrows_per_gpu, d, device="cuda", dtype=tdt, generator=torch.Generator(device="cuda").manual_seed(200 + r)) / d ** 0.5
                 for r in range(W_)]
        X = torch.stack(parts).reshape(W_, nb, batch_per_gpu, d).permute(1, 0, 2, 3).reshape(-1, d).contiguous()
        del parts
        batch_per_gpu, nrows_per_gpu = batch_per_gpu * W_, nrows_per_gpu * W_
    else:
        gen_r = torch.Generator(device="cuda").manual_seed(200 + rank)
        X = torch.randn(nrows_per_gpu, d, device="cuda", dtype=tdt, generator=gen_r) / d ** 0.5
    lab = torch.empty(nrows_per_gpu, device="cuda", dtype=torch.int32)
    for r0 in range(0, nrows_per_gpu, 4096):
        r1 = min(nrows_per_gpu, r0 + 4096)
        lab[r0:r1] = torch.argmax(X[r0:r1] @ Wt.T * 4.0, dim=1).to(torch.int32)
    del Wt
    big = min(nrows_per_gpu, batch_per_gpu * L)
    x_full = torch.zeros(n, device="cuda", dtype=tdt)
    xq = torch.zeros(n, device="cuda", dtype=tdt)              # the requested point, gathered
    g_full = torch.zeros(n, device="cuda", dtype=tdt)
    g_blk = torch.zeros(blk, device="cuda", dtype=tdt) if sharded_opt else g_full
    work = torch.empty(lib.stochqn_b200_multinomial_work_size(max(batch_per_gpu, big), d, K), device="cuda", dtype=torch.uint8)
    sw = {c: torch.full((c,), 1.0 / (c * world), device="cuda", dtype=tdt) for c in {batch_per_gpu, big}}
    alpha = 1e-3
    ws = lib.initialize_adaQN(blk, 10, 1, L, 0.0, 1e-4, 1e-4, rms, 1, 0.0, 1, 1)
    assert ws, _lib.last_error(abi)
    if sharded_opt and comm is not None:
        assert lib.stochqn_b200_set_comm(ws, comm, n) == 0
    x_ptr = x_full.data_ptr() + off * esz
    req, task, info = C.c_void_p(), C.c_int(), C.c_int()
    tasks, infos = {}, {}
    state = dict(b=0)

    def call():
        lib.run_adaQN(step, x_ptr, 0.0, g_blk.data_ptr(), C.byref(req), C.byref(task), ws, C.byref(info))
        tasks[task.value] = tasks.get(task.value, 0) + 1
        infos[info.value] = infos.get(info.value, 0) + 1

    phases = os.environ.get("CFG5S_PHASES", "0") != "0"       # per-phase device times (CUDA events on the stream)
    marks = []

    def mark(tag):
        if phases:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            marks.append((tag, e))

    def serve():
        t = task.value
        b = state["b"]
        if t == 101:
            state["b"] = b = (b + 1) % nb
            r0, cnt = b * batch_per_gpu, batch_per_gpu
        elif t == 103:
            cnt = big
            r0 = 0 if fixed_big else max(0, (b + 1) * batch_per_gpu - cnt)
        else:
            raise RuntimeError("unexpected task %d" % t)
        mark("start")
        _serve(t, b, r0, cnt)
        mark("grad_done")

    def _serve(t, b, r0, cnt):
        if mode in ("fused", "p2p") and world > 1:            # gather by pushing over peer memory (one kernel + barrier)
            gp = C.c_void_p()
            rc = lib.stochqn_b200_all_gather_p2p(comm, req.value, blk, C.byref(gp), None)
            assert rc == 0, (rc, _lib.last_error(abi))
            point = gp.value
            mark("gathered")
        elif sharded_opt:                                     # gather the point the request names (x, x_avg or x_avg_prev block)
            lib.stochqn_b200_all_gather_real(comm, req.value, xq.data_ptr(), blk, None)
            point = xq.data_ptr()
            mark("gathered")
        else:
            point = req.value
        if mode == "fused" and world > 1:
            # gradient on this rank's rows with the reduce-scatter fused into the epilogue of the product that computes it
            rc = lib.stochqn_b200_multinomial_grad_reduce_scatter(comm, X.data_ptr() + r0 * d * esz, d, None, K, lab.data_ptr() + r0 * 4,
                                                                  sw[cnt].data_ptr(), cnt, d, K, 1, point, alpha / world, g_blk.data_ptr(), blk,
                                                                  work.data_ptr(), None)
            assert rc == 0, (rc, _lib.last_error(abi))
            return
        if mode == "p2p" and world > 1:
            # the gradient lands in the library's peer-mapped send vector; the owner of every block then pulls it from all
            # the ranks over NVLink and adds in rank order (rank barrier + one kernel instead of ncclReduceScatter)
            sp = C.c_void_p()
            rc = lib.stochqn_b200_p2p_send_buffer(comm, blk, C.byref(sp))
            assert rc == 0, (rc, _lib.last_error(abi))
            assert lib.stochqn_b200_multinomial_loss_grad(X.data_ptr() + r0 * d * esz, d, None, K, lab.data_ptr() + r0 * 4, sw[cnt].data_ptr(), cnt, d, K, 1,
                                                          point, alpha / world, sp.value, None, work.data_ptr(), None) == 0
            mark("grad_local")
            rc = lib.stochqn_b200_reduce_scatter_p2p(comm, sp.value, g_blk.data_ptr(), blk, None)
            assert rc == 0, (rc, _lib.last_error(abi))
            return
        # alpha / world per rank: the penalty term is added once in the sum over ranks
        assert lib.stochqn_b200_multinomial_loss_grad(X.data_ptr() + r0 * d * esz, d, None, K, lab.data_ptr() + r0 * 4, sw[cnt].data_ptr(), cnt, d, K, 1,
                                                      point, alpha / world, g_full.data_ptr(), None, work.data_ptr(), None) == 0
        if sharded_opt:
            lib.stochqn_b200_reduce_scatter_real(comm, g_full.data_ptr(), g_blk.data_ptr(), blk, None)
        else:
            lib.stochqn_b200_allreduce_real(comm, g_full.data_ptr(), n, None)

    call()
    niter = lambda: int(ws.contents.niter)
    while niter() < warm_cycles * L:
        serve(); call()
    torch.cuda.synchronize()
    if world > 1:
        if mode in ("fused", "p2p"):
            # a peer-memory exchange that gave up waiting (20 s, sticky) leaves garbage behind: every rank learns of it and
            # all of them leave together, so that the caller can fall back to the library collectives
            bad = torch.tensor([float(lib.stochqn_b200_comm_error(comm))], device="cuda")
            dist.all_reduce(bad, op=dist.ReduceOp.MAX)
            if float(bad.item()) != 0.0:
                lib.dealloc_adaQN(ws)
                lib.stochqn_b200_comm_destroy(comm)
                raise RuntimeError("a peer-memory exchange timed out in mode %s" % mode)
        dist.barrier()
        torch.cuda.synchronize()
    tasks.clear(); infos.clear()
    marks.clear()
    it0 = niter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while niter() < it0 + steps:
        serve(); call()
    e1.record()
    mark("start")
    torch.cuda.synchronize()
    phase_ms = {}
    if phases:
        for (ta, ea), (tb, eb) in zip(marks[:-1], marks[1:]):
            key = ta + "->" + tb
            phase_ms[key] = phase_ms.get(key, 0.0) + ea.elapsed_time(eb) / steps
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    if sharded_opt and world > 1:
        lib.stochqn_b200_all_gather_real(comm, x_ptr, x_full.data_ptr(), blk, None)
    torch.cuda.synchronize()
    out = dict(config=name, optimizer="adaQN", mode=mode if world > 1 else "single", n_gpus=world, dtype="f64" if esz == 8 else "f32", n=n,
               features=d, classes=K, batch_per_gpu=batch_per_gpu, global_batch=batch_per_gpu * world, steps=steps, ms_per_step=ms,
               steps_per_s=1e3 / ms, samples_per_s=1e3 / ms * batch_per_gpu * world, tasks=tasks, infos=infos,
               mem_used=int(ws.contents.bfgs_memory.contents.mem_used), x_norm=float(torch.linalg.vector_norm(x_full.double()).item()))
    if phases:
        out["phase_ms_rank0"] = {k: round(v, 4) for k, v in phase_ms.items()}
    lib.dealloc_adaQN(ws)
    if rank == 0 and not quiet:
        print(json.dumps(out), flush=True)
    xr = x_full.cpu().numpy().astype(np.float64) if return_x else None
    if world > 1:
        dist.barrier()
        lib.stochqn_b200_comm_destroy(comm)
    return out, xr


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="*", default=["cfg1", "cfg2", "cfg3", "cfg5"])
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--rows-cfg2", type=int, default=1000000)
    ap.add_argument("--mode", default="zero1", choices=["zero1", "allreduce", "fused", "p2p"],
                    help="cfg5s: how the row-sharded gradient is combined (zero1: ncclReduceScatter / ncclAllGather; p2p: push all-gather + pull "
                         "reduce-scatter over NVLink peer memory; fused: reduce-scatter inside the GEMM epilogue over peer memory)")
    a = ap.parse_args()
    if "cfg5s" in a.configs:          # row-sharded config 5 (torch.distributed.run, one rank per GPU)
        import torch.distributed as dist
        run_multinomial_sharded("cfg5 row-sharded", np.float32, 8192, 4096, 1024, 16384, min(a.steps, 100), 10, 0.9, 1e-3, mode=a.mode)
        if dist.is_initialized():
            dist.destroy_process_group()
        return
    if "cfg1" in a.configs:
        run_logistic("cfg1", "oLBFGS", 100000, 1000, 1000, a.steps)
    if "cfg1n" in a.configs:
        run_logistic("cfg1", "oLBFGS", 100000, 1000, 1000, a.steps, native=True)
    if "cfg1d" in a.configs:
        run_logistic("cfg1", "oLBFGS", 100000, 1000, 1000, a.steps, native="batches")
    if "cfg2d" in a.configs:
        run_logistic("cfg2", "SQN", a.rows_cfg2, 4096, 2000, a.steps, L=10, big=20000, native="batches", step=1e-2)
    if "cfg2" in a.configs:
        run_logistic("cfg2", "SQN", a.rows_cfg2, 4096, 2000, a.steps, L=10, big=20000, step=1e-2)
    if "cfg2n" in a.configs:
        run_logistic("cfg2", "SQN", a.rows_cfg2, 4096, 2000, a.steps, L=10, big=20000, native=True, step=1e-2)
    if "cfg3" in a.configs:
        run_multinomial("cfg3", np.float64, 1836, 159, 50, 6655, a.steps, 20, 100, 0, 1.01, 0.0, 1e-2, profile="bibtex")
    if "cfg3d" in a.configs:
        run_multinomial("cfg3", np.float64, 1836, 159, 50, 6655, a.steps, 20, 100, 0, 1.01, 0.0, 1e-2, profile="bibtex", native=True)
    if "cfg5" in a.configs:
        run_multinomial("cfg5", np.float32, 8192, 4096, 1024, 16384, min(a.steps, 100), 10, 0, 1, 0.0, 0.9, 1e-3, fixed_big=True)


if __name__ == "__main__":
    main()
