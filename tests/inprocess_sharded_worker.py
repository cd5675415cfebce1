"""Worker of tests/test_gpu_inprocess_collectives.py::test_sharded_*: the SHARDED optimizer on ONE GPU.

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
from stochqn_b200.distributed import shard_bounds       # noqa: E402


def run_case(kind, n, calls, world):
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
                call()
        except Exception as e:                             # noqa: BLE001
            errors[r] = "%s: %s\n%s" % (type(e).__name__, e, traceback.format_exc()[-800:])

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
    to = run_trace(so, p, calls, step, keep_x=True)
    want = [(r["task"], r["ret"], r["info"], r["niter"], r["section"], r["mem_used"], r["mem_st_ix"]) for r in to]
    err = float(np.max(np.abs(got - to[-1]["x"])) / np.max(np.abs(to[-1]["x"])))
    res = dict(kind=kind, world=world, rel_err=err, same_on_all_ranks=all(t == traces[0] for t in traces),
               matches_oracle=(traces[0] == want), pairs=int(to[-1]["mem_used"]), exchange_timeouts=timed_out)
    assert res["same_on_all_ranks"] and res["matches_oracle"], res
    assert err <= 1e-10 and res["pairs"] >= 4 and not any(timed_out), res
    return res


def main():
    out = sys.argv[1]
    res = {}
    for kind, world in (("oLBFGS", 2), ("SQN", 2), ("oLBFGS", 4), ("SQN", 3)):
        name = "%s_w%d" % (kind, world)
        try:
            r = run_case(kind, 100003, 90, world)
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
