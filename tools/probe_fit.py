"""Phase timeline of the fused mini-batch kernel (csrc/kernels_fit.cuh): CTA 0 stamps %globaltimer at every phase boundary.

    python tools/probe_fit.py [cfg1|cfg2] [--batches 50]

Prints the median duration of each phase over the traced mini-batches (ns) and the step time they add up to."""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stochqn_b200 import _lib

NAMES = ["R rows", "barrier 1", "pair decision (previous mini-batch) + C columns", "dots", "barrier 2", "S reduce+solve+update", "barrier 3", "R' rows",
         "barrier 4 + C'", "pair partials"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", nargs="?", default="cfg1")
    ap.add_argument("--batches", type=int, default=50)
    a = ap.parse_args()
    kind, d, batch = ("oLBFGS", 1000, 1000) if a.config == "cfg1" else ("SQN", 4096, 2000)
    n = d + 1
    nb = a.batches
    abi = _lib.load(np.float64)
    lib = abi.lib
    g = torch.Generator(device="cuda").manual_seed(1)
    X = torch.randn(batch * nb, n, device="cuda", dtype=torch.float64, generator=g)
    X[:, 0] = 1.0
    w = torch.randn(n, device="cuda", dtype=torch.float64, generator=g) * (2.0 / d ** 0.5)
    y = (torch.rand(batch * nb, device="cuda", dtype=torch.float64, generator=g) < torch.sigmoid(X @ w)).double()
    x = torch.zeros(n, device="cuda", dtype=torch.float64)
    g0 = torch.zeros(n, device="cuda", dtype=torch.float64)
    work = torch.empty(lib.stochqn_b200_logistic_work_size(batch, n), device="cuda", dtype=torch.uint8)
    ws = lib.initialize_oLBFGS(n, 10, 0.0, 0.0, 1e-4, 1, 1) if kind == "oLBFGS" else lib.initialize_SQN(n, 10, 1 << 30, 1e-4, 0, 0.0, 1, 1)
    req, req_vec, task, info = C.c_void_p(), C.c_void_p(), C.c_int(), C.c_int()
    if kind == "oLBFGS":
        lib.run_oLBFGS(0.1, x.data_ptr(), g0.data_ptr(), C.byref(req), C.byref(task), ws, C.byref(info))
    else:
        lib.run_SQN(0.01, x.data_ptr(), g0.data_ptr(), g0.data_ptr(), C.byref(req), C.byref(req_vec), C.byref(task), ws, C.byref(info))
    M = abi.Model(0, 0, n, 0, 1e-5, work.data_ptr())
    data = _lib.Rows(X.data_ptr(), n, y.data_ptr(), 1, None, batch * nb)
    rep = _lib.FitReport()
    trace = torch.zeros(16 * nb, device="cuda", dtype=torch.int64)
    step = 0.1 if kind == "oLBFGS" else 0.01
    for rnd in range(3):          # the last round is the one reported (memory full, everything warm)
        lib.stochqn_b200_debug_fit_trace(ws, trace.data_ptr() if rnd == 2 else None)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.stochqn_b200_fit_batches(ws, x.data_ptr(), step, C.byref(M), C.byref(data), 0, batch, nb, None, None, None, C.byref(task),
                                          C.byref(req), C.byref(req_vec), C.byref(rep))
        e1.record()
        torch.cuda.synchronize()
        assert rc == 0, _lib.last_error(abi)
        ms = e0.elapsed_time(e1)
    t = trace.cpu().numpy().reshape(nb, 16).astype(np.float64)
    last = 10 if kind == "oLBFGS" else 7
    dur = np.diff(t[:, :last + 1], axis=1)
    med = np.median(dur[nb // 4:], axis=0)
    inner = np.median(np.stack([t[:, 12] - t[:, 5], t[:, 13] - t[:, 12], t[:, 14] - t[:, 13], t[:, 6] - t[:, 14]], axis=1)[nb // 4:], axis=0)
    per_batch = np.median(np.diff(t[:, 0]))
    out = {"config": a.config, "optimizer": kind, "n": n, "batch": batch, "us_per_step_events": ms * 1e3 / nb, "us_per_step_trace": per_batch / 1e3,
           "phases_ns": {NAMES[i]: float(med[i]) for i in range(len(med))}, "S_split_ns": {"reduce records": float(inner[0]), "solve": float(inner[1]), "combine": float(inner[2]), "update": float(inner[3])}, "fit_steps": _lib.get_stat(abi, ws, _lib.STAT_FUSED_FIT_STEPS)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
