"""GPU probe: per-kernel device times of the oLBFGS step at several n (dev tool, not a bench)."""
import ctypes as C
import json
import sys
import os
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stochqn_b200 import _lib


def run(n, m=10, iters=int(os.environ.get('PROBE_ITERS', 30)), warm=12, dtype=np.float64, writeback=1):
    abi = _lib.load(dtype)
    lib = abi.lib
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    esz = 8 if dtype == np.float64 else 4
    x = torch.empty(n, device="cuda", dtype=tdt)
    g = torch.empty(n, device="cuda", dtype=tdt)
    lib.stochqn_b200_rosenbrock_x0(x.data_ptr(), n, 0, None)
    ws = lib.initialize_oLBFGS(n, m, 0.0, 0.0, 1e-4, 1, 1)
    assert ws, _lib.last_error(abi)
    lib.stochqn_b200_set_option(ws, _lib.OPT_GRAD_WRITEBACK, writeback)
    req, task, info = C.c_void_p(), C.c_int(), C.c_int()
    step = 1e-4
    lib.run_oLBFGS(step, x.data_ptr(), g.data_ptr(), C.byref(req), C.byref(task), ws, C.byref(info))
    infos = 0

    def iteration():
        nonlocal infos
        for _ in range(2):
            lib.stochqn_b200_rosenbrock_grad(req.value, g.data_ptr(), n, 0, n, None, None)
            lib.run_oLBFGS(step, x.data_ptr(), g.data_ptr(), C.byref(req), C.byref(task), ws, C.byref(info))
            infos += info.value != 200
            if task.value == 101:
                break

    for _ in range(warm):
        iteration()
    lib.stochqn_b200_set_option(ws, _lib.OPT_PROFILE, 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    for _ in range(iters):
        iteration()
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = e0.elapsed_time(e1) / iters
    st = {k: _lib.get_stat(abi, ws, v) for k, v in dict(k1=1, k1n=2, k3=3, k3n=4, k4=5, k4n=6, U=7).items()}
    vec = n * esz
    k1 = st["k1"] / max(st["k1n"], 1); k3 = st["k3"] / max(st["k3n"], 1); k4 = st["k4"] / max(st["k4n"], 1)
    used = ws.contents.bfgs_memory.contents.mem_used
    out = dict(n=n, m=m, dtype=np.dtype(dtype).name, writeback=writeback, ms_per_iter=ms, wall_ms_per_iter=1e3 * wall / iters,
               steps_per_s=1e3 / ms, k1_ms=k1, k3_ms=k3, k4_ms=k4,
               k1_gbs=(2 * used + 2) * vec / k1 / 1e6, k3_gbs=(2 * used + 4) * vec / k3 / 1e6, k4_gbs=4 * vec / k4 / 1e6,
               e2e_gbs=(4 * used + 14) * vec / ms / 1e6, infos=infos, mem_used=int(used), niter=int(ws.contents.niter),
               U=st["U"], xnorm=float(torch.linalg.vector_norm(x).item()))
    # stand-alone timing of the gradient callback and of an empty-ish call sequence
    e0.record()
    for _ in range(20):
        lib.stochqn_b200_rosenbrock_grad(x.data_ptr(), g.data_ptr(), n, 0, n, None, None)
    e1.record()
    torch.cuda.synchronize()
    out["grad_ms"] = e0.elapsed_time(e1) / 20
    out["grad_gbs"] = 2 * vec / out["grad_ms"] / 1e6
    out["overhead_ms"] = ms - (k1 + k3 + k4 + 2 * out["grad_ms"])
    lib.dealloc_oLBFGS(ws)
    del x, g
    torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    sizes = [int(a) for a in sys.argv[1:]] or [2 ** 20, 2 ** 24, 2 ** 27]
    for n in sizes:
        for dt in [d for d, tag in ((np.float64, "f64"), (np.float32, "f32")) if tag in os.environ.get("PROBE_DTYPES", "f64,f32")]:
            r = run(n, dtype=dt)
            print(json.dumps(r), flush=True)
