"""Worker of tests/test_gpu_multi.py: one rank of a sharded optimisation (launched by torch.distributed.run).

Runs free-mode oLBFGS / SQN(grad-diff) on the chained Rosenbrock function with the parameter vector sharded by
contiguous blocks over WORLD_SIZE GPUs, every request served by the bundled device callback (halo exchange over
the library's communicator); rank 0 gathers x and compares it with the oracle run on the whole vector.
argv: kind n calls out_json
"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

from oracle import stochqn_np as O                      # noqa: E402
from oracle.driver import HostStepper, discrete, run_trace  # noqa: E402
from oracle.problems import Rosenbrock                  # noqa: E402
from stochqn_b200 import _lib                           # noqa: E402
from stochqn_b200.distributed import init_comm, shard_bounds  # noqa: E402


def main():
    kind, n, calls, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    abi = _lib.load(np.float64)
    lib = abi.lib
    off, cnt = shard_bounds(n, rank, world)
    comm = init_comm(abi, rank, world)
    uses_p2p = int(lib.stochqn_b200_comm_uses_p2p(comm))
    x = torch.empty(cnt, device="cuda", dtype=torch.float64)
    g = torch.zeros(cnt, device="cuda", dtype=torch.float64)
    halo = torch.zeros(2, device="cuda", dtype=torch.float64)
    scratch = torch.zeros(2 * world, device="cuda", dtype=torch.float64)
    lib.stochqn_b200_rosenbrock_x0(x.data_ptr(), cnt, off, None)
    step = 1e-4
    if kind == "oLBFGS":
        kw = dict(mem_size=5, hess_init=0.0, y_reg=0.0, min_curvature=1e-4, check_nan=1)
        ws = lib.initialize_oLBFGS(cnt, 5, 0.0, 0.0, 1e-4, 1, 1)
    else:
        kw = dict(mem_size=4, bfgs_upd_freq=3, min_curvature=1e-4, use_grad_diff=1, y_reg=0.0, check_nan=1)
        ws = lib.initialize_SQN(cnt, 4, 3, 1e-4, 1, 0.0, 1, 1)
    assert ws, _lib.last_error(abi)
    assert lib.stochqn_b200_set_comm(ws, comm, n) == 0
    req, req_vec, task, info = C.c_void_p(), C.c_void_p(), C.c_int(), C.c_int()
    trace = []

    def call():
        if kind == "oLBFGS":
            ret = lib.run_oLBFGS(step, x.data_ptr(), g.data_ptr(), C.byref(req), C.byref(task), ws, C.byref(info))
        else:
            ret = lib.run_SQN(step, x.data_ptr(), g.data_ptr(), None, C.byref(req), C.byref(req_vec), C.byref(task), ws, C.byref(info))
        w = ws.contents
        m = w.bfgs_memory.contents
        trace.append((int(task.value), int(ret), int(info.value), int(w.niter), int(w.section), int(m.mem_used), int(m.mem_st_ix)))

    call()
    for _ in range(calls - 1):
        assert task.value in (101, 102, 103), task.value
        if len(trace) % 2:      # alternate between the fused-halo gradient kernel and the two-step form
            lib.stochqn_b200_rosenbrock_grad_sharded(req.value, g.data_ptr(), cnt, off, n, rank, world, comm, halo.data_ptr(),
                                                     scratch.data_ptr(), None)
        else:
            lib.stochqn_b200_rosenbrock_halo(req.value, cnt, rank, world, comm, halo.data_ptr(), scratch.data_ptr(), None)
            lib.stochqn_b200_rosenbrock_grad(req.value, g.data_ptr(), cnt, off, n, halo.data_ptr(), None)
        call()
    torch.cuda.synchronize()
    parts = [torch.empty(shard_bounds(n, r, world)[1], device="cuda", dtype=torch.float64) for r in range(world)]
    dist.all_gather(parts, x)
    xs = torch.cat(parts).cpu().numpy()
    traces = [None] * world
    dist.all_gather_object(traces, trace)
    {"oLBFGS": lib.dealloc_oLBFGS, "SQN": lib.dealloc_SQN}[kind](ws)
    if rank == 0:
        p = Rosenbrock(n)
        so = HostStepper({"oLBFGS": O.OracleOLBFGS, "SQN": O.OracleSQN}[kind](n, **kw), p.x0())
        to = run_trace(so, p, calls, step, keep_x=True)
        want = [(r["task"], r["ret"], r["info"], r["niter"], r["section"], r["mem_used"], r["mem_st_ix"]) for r in to]
        err = float(np.max(np.abs(xs - to[-1]["x"])) / np.max(np.abs(to[-1]["x"])))
        res = dict(kind=kind, world=world, uses_p2p=uses_p2p, rel_err=err,
                   same_on_all_ranks=all(t == traces[0] for t in traces), matches_oracle=(traces[0] == want),
                   pairs=int(to[-1]["mem_used"]))
        json.dump(res, open(out, "w"))
    dist.barrier()
    lib.stochqn_b200_comm_destroy(comm)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
