"""Worker of tests/test_gpu_zz_inprocess.py::test_sharded_*: the SHARDED optimizer on ONE GPU.

The parameter vector is split into contiguous blocks over the ranks of an in-process communicator group
(`stochqn_b200_comm_init_inprocess`); every rank is a host thread with its own workspace and its own stream, serving the
requests with the bundled sharded Rosenbrock callback.  The kernels are those of the one-process-per-GPU case: the
exchange of the 4m+2 partial sums fused into K2, the two curvature dots exchanged by the last CTA of K4, the halo
exchange fused into the gradient kernel - only the mapping of the peers' mailboxes differs (pointers instead of cudaIpc).
Task / counter sequences must be identical on all ranks and equal to the oracle's on the whole vector, iterates 1e-10.

Launched with CUDA_MODULE_LOADING=EAGER and CUDA_DEVICE_MAX_CONNECTIONS=32 (see inprocess_collectives_worker.py).
argv: out_json
"""
import ctypes as C
import json
import os
import sys
import threading
import traceback

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

from oracle import stochqn_np as O                      # noqa: E402
from oracle.driver import HostStepper, run_trace        # noqa: E402
from oracle.problems import Rosenbrock                  # noqa: E402
from stochqn_b200 import _lib                           # noqa: E402
from stochqn_b200.distributed import RowShardedCombiner, shard_bounds       # noqa: E402


def run_case(kind, n, calls, world, events=None):
    """events: {call index: ("remember",) | ("repeat",) | ("poison", global element, value)} - forced events, applied to the
    gradient that is handed to that call on every rank / on the rank that owns the element, and to the oracle's."""
    events = events or {}
    abi = _lib.load(np.float64)
    lib = abi.lib
    arr = (C.c_void_p * world)()
    rc = lib.stochqn_b200_comm_init_inprocess(world, arr)
    assert rc == 0, (rc, _lib.last_error(abi))
    comms = [C.c_void_p(arr[r]) for r in range(world)]
    streams = [torch.cuda.Stream() for _ in range(world)]
    bounds = [shard_bounds(n, r, world) for r in range(world)]
    x_full = torch.empty(n, device="cuda", dtype=torch.float64)
    xs = [x_full[o:o + c] for o, c in bounds]
    gs = [torch.zeros(c, device="cuda", dtype=torch.float64) for _, c in bounds]
    halos = [torch.zeros(2, device="cuda", dtype=torch.float64) for _ in range(world)]
    scratch = [torch.zeros(2 * world, device="cuda", dtype=torch.float64) for _ in range(world)]
    step = 1e-4
    if kind == "oLBFGS":
        kw = dict(mem_size=5, hess_init=0.0, y_reg=0.0, min_curvature=1e-4, check_nan=1)
    else:
        kw = dict(mem_size=4, bfgs_upd_freq=3, min_curvature=1e-4, use_grad_diff=1, y_reg=0.0, check_nan=1)
    wss = []
    for r, (off, cnt) in enumerate(bounds):
        lib.stochqn_b200_rosenbrock_x0(xs[r].data_ptr(), cnt, off, None)
        ws = lib.initialize_oLBFGS(cnt, 5, 0.0, 0.0, 1e-4, 1, 1) if kind == "oLBFGS" else lib.initialize_SQN(cnt, 4, 3, 1e-4, 1, 0.0, 1, 1)
        assert ws, _lib.last_error(abi)
        assert lib.stochqn_b200_set_stream(ws, C.c_void_p(streams[r].cuda_stream)) == 0
        assert lib.stochqn_b200_set_comm(ws, comms[r], n) == 0
        wss.append(ws)
    torch.cuda.synchronize()
    traces = [[] for _ in range(world)]
    errors = [None] * world
    start = threading.Barrier(world)

    def rank_main(r):
        try:
            off, cnt = bounds[r]
            ws, comm, st = wss[r], comms[r], C.c_void_p(streams[r].cuda_stream)
            x, g, halo, scr = xs[r].data_ptr(), gs[r].data_ptr(), halos[r].data_ptr(), scratch[r].data_ptr()
            req, req_vec, task, info = C.c_void_p(), C.c_void_p(), C.c_int(), C.c_int()
            trace = traces[r]

            def call():
                if kind == "oLBFGS":
                    ret = lib.run_oLBFGS(step, x, g, C.byref(req), C.byref(task), ws, C.byref(info))
                else:
                    ret = lib.run_SQN(step, x, g, None, C.byref(req), C.byref(req_vec), C.byref(task), ws, C.byref(info))
                w = ws.contents
                m = w.bfgs_memory.contents
                trace.append((int(task.value), int(ret), int(info.value), int(w.niter), int(w.section), int(m.mem_used), int(m.mem_st_ix)))

            start.wait(timeout=60)
            call()
            for _ in range(calls - 1):
                assert task.value in (101, 102, 103), task.value
                if len(trace) % 2:      # alternate between the fused-halo gradient kernel and the two-step form
                    rc = lib.stochqn_b200_rosenbrock_grad_sharded(req.value, g, cnt, off, n, r, world, comm, halo, scr, st)
                else:
                    rc = lib.stochqn_b200_rosenbrock_halo(req.value, cnt, r, world, comm, halo, scr, st)
                    rc = rc or lib.stochqn_b200_rosenbrock_grad(req.value, g, cnt, off, n, halo, st)
                assert rc == 0, (rc, _lib.last_error(abi))
                ev = events.get(len(trace))
                if ev:
                    with torch.cuda.stream(streams[r]):    # behind the gradient kernel on the rank's stream
                        if ev[0] == "remember":
                            kept[r].copy_(gs[r])
                        elif ev[0] == "repeat":            # the same gradient again: y = 0, the pair is rejected (quirk Q1)
                            gs[r].copy_(kept[r])
                        elif ev[0] == "poison" and off <= ev[1] < off + cnt:
                            gs[r][ev[1] - off] = ev[2]
                call()
        except Exception as e:                             # noqa: BLE001
            errors[r] = "%s: %s\n%s" % (type(e).__name__, e, traceback.format_exc()[-800:])

    kept = [torch.zeros_like(g) for g in gs]
    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=240)
    hung = [r for r, t in enumerate(threads) if t.is_alive()]
    assert not hung, "ranks %s did not finish" % hung
    assert all(e is None for e in errors), errors
    torch.cuda.synchronize()
    timed_out = [int(lib.stochqn_b200_comm_error(c)) for c in comms]
    got = x_full.cpu().numpy()
    for ws in wss:
        {"oLBFGS": lib.dealloc_oLBFGS, "SQN": lib.dealloc_SQN}[kind](ws)
    for c in comms:
        lib.stochqn_b200_comm_destroy(c)
    p = Rosenbrock(n)
    so = HostStepper({"oLBFGS": O.OracleOLBFGS, "SQN": O.OracleSQN}[kind](n, **kw), p.x0())
    held = {}

    def hook(ev):
        def f(stepper, task, payload):
            if ev[0] == "remember":
                held["g"] = payload["grad"].copy()
            elif ev[0] == "repeat":
                payload["grad"] = held["g"].copy()
            else:
                payload["grad"] = payload["grad"].copy()
                payload["grad"][ev[1]] = ev[2]
        return f

    to = run_trace(so, p, calls, step, hooks={c: hook(ev) for c, ev in events.items()}, keep_x=True)
    want = [(r["task"], r["ret"], r["info"], r["niter"], r["section"], r["mem_used"], r["mem_st_ix"]) for r in to]
    err = float(np.max(np.abs(got - to[-1]["x"])) / np.max(np.abs(to[-1]["x"])))
    infos = sorted(set(t[2] for t in traces[0]))
    res = dict(kind=kind, world=world, rel_err=err, same_on_all_ranks=all(t == traces[0] for t in traces),
               matches_oracle=(traces[0] == want), pairs=int(to[-1]["mem_used"]), exchange_timeouts=timed_out, infos=infos)
    if not res["matches_oracle"]:
        res["first_difference"] = next(((i, a, b) for i, (a, b) in enumerate(zip(traces[0], want)) if a != b), None)
    assert res["same_on_all_ranks"] and res["matches_oracle"], res
    assert err <= 1e-10 and not any(timed_out), res
    if events:
        assert 203 in infos, res                                   # a step was rejected and the memory flushed ...
        if any(e[0] == "repeat" for e in events.values()):
            assert 202 in infos, res                               # ... and a pair was rejected
    else:
        assert res["pairs"] >= 4, res
    return res


def run_rowsharded_adaqn(world, d=64, K=40, bpg=160, nb=10, steps=45, L=5, rms=0.9, step=1e-2):
    """adaQN (RMSProp, gradient-difference pairs) + multinomial gradient with the batch ROWS sharded over in-process ranks
    and the optimizer state sharded by blocks: per request push all-gather of the point, gradient on the rank's rows into
    the library's send vector, pull reduce-scatter, sharded run_adaQN (BASELINE config 5's multi-GPU layout at a small
    size, fp64) - against the same optimisation of the union of the rows on one rank."""
    abi = _lib.load(np.float64)
    lib = abi.lib
    tdt, esz = torch.float64, 8
    n = K * (d + 1)
    assert n % world == 0
    blk = n // world
    gen = torch.Generator(device="cuda").manual_seed(11)
    Wt = torch.randn(K, d, device="cuda", dtype=tdt, generator=gen)
    Xu4 = torch.randn(nb, world, bpg, d, device="cuda", dtype=tdt, generator=gen) / d ** 0.5      # [batch][rank][row]
    Xu = Xu4.reshape(-1, d).contiguous()
    labu = torch.argmax(Xu @ Wt.T * 4.0, dim=1).to(torch.int32)
    Xr = [Xu4[:, r].reshape(-1, d).contiguous() for r in range(world)]
    labr = [labu.reshape(nb, world, bpg)[:, r].reshape(-1).contiguous() for r in range(world)]
    alpha = 1e-3
    big = bpg * L                                        # rows per rank of the long batch (the last L batches)

    def optimise(nranks, comms, streams, X, lab, nrows_batch, x_full):
        """every rank: the request loop of tools/bench_configs.py run_multinomial_sharded (mode p2p), as a host thread"""
        cnts = {nrows_batch, nrows_batch * L}
        sw = {c: torch.full((c,), 1.0 / (c * nranks), device="cuda", dtype=tdt) for c in cnts}
        work = [torch.empty(lib.stochqn_b200_multinomial_work_size(nrows_batch * L, d, K), device="cuda", dtype=torch.uint8) for _ in range(nranks)]
        b_loc = n // nranks
        gblk = [torch.zeros(b_loc, device="cuda", dtype=tdt) for _ in range(nranks)]
        wss = []
        for r in range(nranks):
            ws = lib.initialize_adaQN(b_loc, 10, 1, L, 0.0, 1e-4, 1e-4, rms, 1, 0.0, 1, 1)
            assert ws, _lib.last_error(abi)
            if nranks > 1:
                assert lib.stochqn_b200_set_stream(ws, C.c_void_p(streams[r].cuda_stream)) == 0
                assert lib.stochqn_b200_set_comm(ws, comms[r], n) == 0
            wss.append(ws)
        if nranks > 1:
            # the first use of a collective allocates for the whole group and drains the device: done here, rank after rank in
            # two passes, before the rank threads exist
            for r in range(nranks):
                gp = C.c_void_p()
                assert lib.stochqn_b200_all_gather_p2p(comms[r], gblk[r].data_ptr(), b_loc, C.byref(gp), C.c_void_p(streams[r].cuda_stream)) == 0
            for r in range(nranks):
                sp = C.c_void_p()
                assert lib.stochqn_b200_p2p_send_buffer(comms[r], b_loc, C.byref(sp)) == 0
                assert lib.stochqn_b200_reduce_scatter_p2p(comms[r], sp.value, gblk[r].data_ptr(), b_loc, C.c_void_p(streams[r].cuda_stream)) == 0
        comb = [RowShardedCombiner(abi, comms[r], n, nranks, stream=C.c_void_p(streams[r].cuda_stream)) for r in range(nranks)] if nranks > 1 else None
        assert comb is None or all(c.p2p for c in comb)
        torch.cuda.synchronize()
        logs = [dict(tasks={}, infos={}) for _ in range(nranks)]
        errors = [None] * nranks

        def rank_main(r):
            try:
                ws = wss[r]
                st = C.c_void_p(streams[r].cuda_stream) if nranks > 1 else None
                x_ptr = x_full.data_ptr() + r * b_loc * esz
                req, task, info = C.c_void_p(), C.c_int(), C.c_int()
                b = 0

                def call():
                    lib.run_adaQN(step, x_ptr, 0.0, gblk[r].data_ptr(), C.byref(req), C.byref(task), ws, C.byref(info))
                    logs[r]["tasks"][task.value] = logs[r]["tasks"].get(task.value, 0) + 1
                    logs[r]["infos"][info.value] = logs[r]["infos"].get(info.value, 0) + 1

                call()
                while int(ws.contents.niter) < steps:
                    t = task.value
                    if t == 101:
                        b = (b + 1) % nb
                        r0, cnt = b * nrows_batch, nrows_batch
                    elif t == 103:
                        cnt = nrows_batch * L
                        r0 = max(0, (b + 1) * nrows_batch - cnt)
                    else:
                        raise RuntimeError("unexpected task %d" % t)
                    args = (X[r].data_ptr() + r0 * d * esz, d, None, K, lab[r].data_ptr() + r0 * 4, sw[cnt].data_ptr(), cnt, d, K, 1)
                    if nranks == 1:
                        rc = lib.stochqn_b200_multinomial_loss_grad(*args, req.value, alpha, gblk[0].data_ptr(), None, work[0].data_ptr(), None)
                        assert rc == 0, (rc, _lib.last_error(abi))
                    else:                                  # the product-level helper (stochqn_b200.distributed)
                        point = comb[r].gather(req.value)
                        send = comb[r].send_buffer()
                        rc = lib.stochqn_b200_multinomial_loss_grad(*args, point, alpha / nranks, send, None, work[r].data_ptr(), st)
                        assert rc == 0, (rc, _lib.last_error(abi))
                        comb[r].reduce_scatter(gblk[r].data_ptr())
                    call()
            except Exception as e:                         # noqa: BLE001
                errors[r] = "%s: %s\n%s" % (type(e).__name__, e, traceback.format_exc()[-800:])

        if nranks == 1:
            rank_main(0)
        else:
            threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(nranks)]
            for t in threads:
                t.start()
            for t in threads:
                t.join(timeout=240)
            hung = [r for r, t in enumerate(threads) if t.is_alive()]
            assert not hung, "ranks %s did not finish" % hung
        assert all(e is None for e in errors), errors
        torch.cuda.synchronize()
        used = int(wss[0].contents.bfgs_memory.contents.mem_used)
        for ws in wss:
            lib.dealloc_adaQN(ws)
        return logs, used

    # sharded
    arr = (C.c_void_p * world)()
    assert lib.stochqn_b200_comm_init_inprocess(world, arr) == 0, _lib.last_error(abi)
    comms = [C.c_void_p(arr[r]) for r in range(world)]
    streams = [torch.cuda.Stream() for _ in range(world)]
    x_sh = torch.zeros(n, device="cuda", dtype=tdt)
    logs, used = optimise(world, comms, streams, Xr, labr, bpg, x_sh)
    timed_out = [int(lib.stochqn_b200_comm_error(c)) for c in comms]
    for c in comms:
        lib.stochqn_b200_comm_destroy(c)
    # the union of the rows on one rank
    x_un = torch.zeros(n, device="cuda", dtype=tdt)
    logu, used_u = optimise(1, None, None, [Xu], [labu], bpg * world, x_un)
    a, bb = x_sh.cpu().numpy(), x_un.cpu().numpy()
    err = float(np.max(np.abs(a - bb)) / np.max(np.abs(bb)))
    res = dict(world=world, rel_err=err, moved=float(np.max(np.abs(bb))), pairs=used, pairs_unsharded=used_u, exchange_timeouts=timed_out,
               same_on_all_ranks=all(l == logs[0] for l in logs), matches_unsharded=(logs[0] == logu[0]), tasks=logs[0]["tasks"])
    assert res["same_on_all_ranks"] and res["matches_unsharded"] and not any(timed_out), res
    assert err <= 1e-9 and res["moved"] > 1e-3 and used == used_u and used >= 2, res
    return res


def main():
    out = sys.argv[1]
    res = {}
    cases = [("%s_w%d" % (kind, world), run_case, (kind, 100003, 90, world)) for kind, world in (("oLBFGS", 2), ("SQN", 2), ("oLBFGS", 4), ("SQN", 3))]
    cases += [("rowsharded_adaQN_w%d" % world, run_rowsharded_adaqn, (world,)) for world in (2, 4)]
    # forced events with the vector sharded: a repeated gradient (y = 0: pair rejected, slot zeroed), then a NaN / an Inf /
    # a huge finite entry on ONE rank's shard - every rank must reject that step, flush and carry on exactly as the oracle
    ev = lambda v: {7: ("remember",), 8: ("repeat",), 21: ("poison", 70001, v)}       # noqa: E731
    cases += [("events_oLBFGS_w2_nan", run_case, ("oLBFGS", 100003, 40, 2, ev(float("nan")))),
              ("events_oLBFGS_w4_huge", run_case, ("oLBFGS", 100003, 40, 4, ev(1e14))),
              ("events_SQN_w3_inf", run_case, ("SQN", 100003, 40, 3, {15: ("poison", 99999, float("inf"))}))]
    for name, fn, a in cases:
        try:
            r = fn(*a)
            r["ok"] = True
        except Exception as e:                             # noqa: BLE001
            r = {"ok": False, "error": "%s: %s" % (type(e).__name__, e), "trace": traceback.format_exc()[-1500:]}
        res[name] = r
        json.dump(res, open(out, "w"), indent=1)
        if not r["ok"]:
            break                                          # a wedged exchange leaves threads / kernels behind: do not pile more on top
    os._exit(0)                                            # daemon-less threads of a failed case must not keep the process alive


if __name__ == "__main__":
    main()
