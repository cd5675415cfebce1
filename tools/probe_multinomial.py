"""GPU probe: multinomial-logistic gradient / Hessian-vector callbacks at the BASELINE shapes.
config 3: batch 50 x 1836 features x 159 classes (fp64, CUDA cores, latency-bound);
config 5: batch 1024 (per GPU) x 8192 features x 4096 classes (fp32; tcgen05 tf32 tensor cores vs CUDA cores)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stochqn_b200 import _lib

CASES = [("cfg3", np.float64, 50, 1836, 159), ("cfg3_f32", np.float32, 50, 1836, 159), ("cfg5", np.float32, 1024, 8192, 4096),
         ("cfg5_b4096", np.float32, 4096, 8192, 4096)]


def run(name, dtype, B, d, K, reps=20):
    abi = _lib.load(dtype)
    lib = abi.lib
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    X = torch.randn(B, d, device="cuda", dtype=tdt) / d ** 0.5
    lab = torch.randint(0, K, (B,), device="cuda", dtype=torch.int32)
    w = torch.randn(K * (d + 1), device="cuda", dtype=tdt) * 0.1
    v = torch.randn(K * (d + 1), device="cuda", dtype=tdt)
    g = torch.empty_like(w)
    loss = torch.zeros(1, device="cuda", dtype=torch.float64)
    work = torch.empty(lib.stochqn_b200_multinomial_work_size(B, d, K), device="cuda", dtype=torch.uint8)
    out = dict(case=name, dtype=np.dtype(dtype).name, B=B, d=d, K=K, n=K * (d + 1))
    for mode in ("tensor", "cuda_cores"):
        if mode == "cuda_cores":
            os.environ["STOCHQN_B200_NO_TENSOR_CORES"] = "1"
        else:
            os.environ.pop("STOCHQN_B200_NO_TENSOR_CORES", None)
        if mode == "tensor" and dtype == np.float64:
            continue
        for kind in ("grad", "hess_vec"):
            def call():
                if kind == "grad":
                    return lib.stochqn_b200_multinomial_loss_grad(X.data_ptr(), d, None, K, lab.data_ptr(), None, B, d, K, 1, w.data_ptr(), 1e-3,
                                                                  g.data_ptr(), loss.data_ptr(), work.data_ptr(), None)
                return lib.stochqn_b200_multinomial_hess_vec(X.data_ptr(), d, None, K, lab.data_ptr(), None, B, d, K, 1, w.data_ptr(), v.data_ptr(),
                                                             1e-3, g.data_ptr(), work.data_ptr(), None)
            r = reps if (mode == "tensor" or B * d * K < 1e9) else 3
            for _ in range(2):
                assert call() == 0
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(r):
                call()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / r
            gemms = 2 if kind == "grad" else 3
            out["%s_%s_ms" % (mode, kind)] = ms
            out["%s_%s_tflops" % (mode, kind)] = gemms * 2.0 * B * d * K / ms / 1e9
    os.environ.pop("STOCHQN_B200_NO_TENSOR_CORES", None)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    sel = sys.argv[1:]
    for c in CASES:
        if not sel or c[0] in sel:
            run(*c)
