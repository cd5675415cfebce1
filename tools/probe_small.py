"""GPU probe: oLBFGS / SQN iterations per second at latency-bound sizes, step taken as K1 -> K2 -> K3 ("three")
or as the fused one-launch kernel of kernels_small.cuh ("one").  Finds the crossover that sets the library's default
threshold (kSmallNDefault).  Dev tool, not a bench.

    python tools/probe_small.py [n ...]
"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stochqn_b200 import _lib


def run(kind, n, route, m=10, iters=300, warm=30, dtype=np.float64):
    abi = _lib.load(dtype)
    lib = abi.lib
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    x = torch.empty(n, device="cuda", dtype=tdt)
    g = torch.empty(n, device="cuda", dtype=tdt)
    hv = torch.empty(n, device="cuda", dtype=tdt)
    lib.stochqn_b200_rosenbrock_x0(x.data_ptr(), n, 0, None)
    if kind == "oLBFGS":
        ws = lib.initialize_oLBFGS(n, m, 0.0, 0.0, 1e-4, 1, 1)
    else:
        ws = lib.initialize_SQN(n, m, 5, 1e-4, 1, 0.0, 1, 1)       # grad-diff pairs every 5 steps
    assert ws, _lib.last_error(abi)
    assert lib.stochqn_b200_set_option(ws, _lib.OPT_ONE_LAUNCH_MAX_N, 0 if route == "three" else 1 << 40) == 0
    req, req_vec, task, info = C.c_void_p(), C.c_void_p(), C.c_int(), C.c_int()
    step = 1e-4

    def call():
        if kind == "oLBFGS":
            lib.run_oLBFGS(step, x.data_ptr(), g.data_ptr(), C.byref(req), C.byref(task), ws, C.byref(info))
        else:
            lib.run_SQN(step, x.data_ptr(), g.data_ptr(), hv.data_ptr(), C.byref(req), C.byref(req_vec), C.byref(task), ws, C.byref(info))

    def until(target):
        while ws.contents.niter < target:
            lib.stochqn_b200_rosenbrock_grad(req.value, g.data_ptr(), n, 0, n, None, None)
            call()

    call()
    until(warm)
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    until(warm + iters)
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / iters
    out = dict(kind=kind, n=n, route=route, dtype=np.dtype(dtype).name, us_per_iter=us, iters_per_s=1e6 / us,
               launches_per_iter=(_lib.launch_count() - l0) / iters,
               one_launch_steps=_lib.get_stat(abi, ws, _lib.STAT_ONE_LAUNCH_STEPS),
               mem_used=int(ws.contents.bfgs_memory.contents.mem_used), xnorm=float(torch.linalg.vector_norm(x).item()))
    (lib.dealloc_oLBFGS if kind == "oLBFGS" else lib.dealloc_SQN)(ws)
    return out


if __name__ == "__main__":
    sizes = [int(a) for a in sys.argv[1:]] or [1001, 4097, 16384, 32768, 65536, 131072, 262144, 1048576]
    for kind in ("oLBFGS", "SQN"):
        for n in sizes:
            for route in ("three", "one"):
                print(json.dumps(run(kind, n, route)), flush=True)
