"""Phase timelines of the two kernels of an ordinary adaQN step at the BASELINE config-3 shape (multinomial 50 x 1836 x 159,
n = 292 083, fp64): mn_grad_small (csrc/multinomial.cu) and kl_ada (csrc/kernels_loop.cuh).  CTA 0 stamps %globaltimer.

    python tools/probe_cfg3.py
"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stochqn_b200 import _lib

MN = ["chunk load + partial Z", "barrier 1", "row sums + softmax", "barrier 2", "load D", "gradient chunk"]
ADA = ["A: G, Fisher row, S'g", "barrier 1", "reduce records", "stage + solve u", "B: Y'[h(Yu-g)]", "barrier 2", "reduce + solve a", "C: combine + update"]


def main():
    d, K, batch, nrows, L = 1836, 159, 50, 6655, 20
    n = K * (d + 1)
    abi = _lib.load(np.float64)
    lib = abi.lib
    tdt = torch.float64
    gen = torch.Generator(device="cuda").manual_seed(4)
    X = (torch.rand(nrows, d, device="cuda", dtype=tdt, generator=gen) < 0.0375).to(tdt)
    lab = torch.randint(0, K, (nrows,), device="cuda", generator=gen)
    Y = torch.zeros(nrows, K, device="cuda", dtype=tdt)
    Y[torch.arange(nrows, device="cuda"), lab] = 1.0
    sw = torch.ones(nrows, device="cuda", dtype=tdt)
    x = torch.randn(n, device="cuda", dtype=tdt, generator=gen)
    g0 = torch.zeros(n, device="cuda", dtype=tdt)
    work = torch.empty(lib.stochqn_b200_multinomial_work_size(batch * L, d, K), device="cuda", dtype=torch.uint8)
    ws = lib.initialize_adaQN(n, 10, 100, L, 0.0, 1e-4, 1e-4, 0.0, 0, 0.0, 1, 1)
    req, req_vec, task, info = C.c_void_p(), C.c_void_p(), C.c_int(), C.c_int()
    lib.run_adaQN(1e-2, x.data_ptr(), 0.0, g0.data_ptr(), C.byref(req), C.byref(task), ws, C.byref(info))
    M = abi.Model(2, 1, d, K, 0.1, work.data_ptr())
    data = _lib.Rows(X.data_ptr(), d, Y.data_ptr(), K, sw.data_ptr(), nrows)
    rep = _lib.FitReport()
    nb = nrows // batch
    tr_mn = torch.zeros(8, device="cuda", dtype=torch.int64)
    tr_ada = torch.zeros(16, device="cuda", dtype=torch.int64)
    mn_d, ada_d, step_us = [], [], []
    b0 = 0
    assert lib.stochqn_b200_set_option(ws, _lib.OPT_FUSED_FIT, 0) == 0      # two launches per step: each kernel is traced on its own
    for it in range(16 * L):
        cnt = 1
        LL = C.c_longlong * cnt
        lf, lr = LL(), LL()
        e = (b0 + 1) * batch
        lr[0] = min(batch * L, e)
        lf[0] = e - lr[0]
        traced = it >= 12 * L and (it + 1) % L != 0
        lib.stochqn_b200_debug_mn_trace(tr_mn.data_ptr() if traced else None)
        lib.stochqn_b200_debug_fit_trace(ws, tr_ada.data_ptr() if traced else None)
        rc = lib.stochqn_b200_fit_batches(ws, x.data_ptr(), 1e-2, C.byref(M), C.byref(data), b0 * batch, batch, cnt, lf, lr, None, C.byref(task),
                                          C.byref(req), C.byref(req_vec), C.byref(rep))
        assert rc == 0, _lib.last_error(abi)
        b0 = (b0 + 1) % nb
        if traced:
            torch.cuda.synchronize()
            a = tr_mn.cpu().numpy().astype(np.float64)
            c = tr_ada.cpu().numpy().astype(np.float64)
            mn_d.append(np.diff(a[:7]))
            ada_d.append(np.diff(c[:9]))
            step_us.append((c[8] - a[0]) / 1e3)
    lib.stochqn_b200_debug_mn_trace(None)
    mn_m, ada_m = np.median(np.array(mn_d), axis=0), np.median(np.array(ada_d), axis=0)
    print(json.dumps({"shape": [batch, d, K], "n": n, "mem_used": int(ws.contents.bfgs_memory.contents.mem_used),
                      "mn_grad_small_ns": {MN[i]: float(mn_m[i]) for i in range(6)}, "mn_total_us": float(mn_m.sum() / 1e3),
                      "kl_ada_ns": {ADA[i]: float(ada_m[i]) for i in range(8)}, "ada_total_us": float(ada_m.sum() / 1e3),
                      "gradient start -> step end, us (median)": float(np.median(step_us))}))
    lib.dealloc_adaQN(ws)


if __name__ == "__main__":
    main()
