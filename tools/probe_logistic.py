"""GPU probe: the bundled binary-logistic callbacks (R/logistic.R:12-37) at the BASELINE shapes, fused one-sweep
kernel against the first two-sweep version.  Algorithmic bytes = one read of the batch (nrows * ncols * sizeof)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stochqn_b200 import _lib

SHAPES = [("cfg1 batch 1000x1001", 1000, 1001), ("cfg2 batch 2000x4097", 2000, 4097), ("cfg2 big batch 20000x4097", 20000, 4097),
          ("100000x1001", 100000, 1001), ("50000x4097", 50000, 4097)]


def run(dtype, only=None):
    abi = _lib.load(dtype)
    lib = abi.lib
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    esz = 8 if dtype == np.float64 else 4
    for name, B, d in SHAPES:
        if only and only not in name:
            continue
        nbatches = max(2, min(64, int(3e9 // (B * d * esz))))        # rotate over > L2 worth of batches
        X = torch.randn(nbatches * B, d, device="cuda", dtype=tdt) / d ** 0.5
        y = (torch.rand(nbatches * B, device="cuda") < 0.5).to(tdt)
        w = torch.randn(d, device="cuda", dtype=tdt)
        v = torch.randn(d, device="cuda", dtype=tdt)
        out = torch.empty(d, device="cuda", dtype=tdt)
        work = torch.empty(lib.stochqn_b200_logistic_work_size(B, d), device="cuda", dtype=torch.uint8)
        res = {}
        for mode in ("fused", "two_sweep"):
            if mode == "two_sweep":
                os.environ["STOCHQN_B200_LOGISTIC_TWO_SWEEP"] = "1"
            else:
                os.environ.pop("STOCHQN_B200_LOGISTIC_TWO_SWEEP", None)
            for kind in ("grad", "hess_vec"):
                def call(i):
                    b = i % nbatches
                    xp = X.data_ptr() + b * B * d * esz
                    yp = y.data_ptr() + b * B * esz
                    if kind == "grad":
                        lib.stochqn_b200_logistic_grad(xp, d, yp, None, B, d, w.data_ptr(), 1e-5, out.data_ptr(), work.data_ptr(), None)
                    else:
                        lib.stochqn_b200_logistic_hess_vec(xp, d, yp, None, B, d, w.data_ptr(), v.data_ptr(), 1e-5, out.data_ptr(), work.data_ptr(), None)
                for i in range(5):
                    call(i)
                torch.cuda.synchronize()
                reps = 40
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(reps):
                    call(i)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                res["%s_%s_ms" % (mode, kind)] = ms
                res["%s_%s_gbs" % (mode, kind)] = B * d * esz / ms / 1e6
        os.environ.pop("STOCHQN_B200_LOGISTIC_TWO_SWEEP", None)
        print(json.dumps(dict(shape=name, dtype=np.dtype(dtype).name, batch_mb=B * d * esz / 1e6, batches_rotated=nbatches, **res)), flush=True)
        del X, y


if __name__ == "__main__":
    only = sys.argv[1] if len(sys.argv) > 1 else None            # substring of a shape name
    dts = [np.float64] if (len(sys.argv) > 2 and sys.argv[2] == "f64") else [np.float64, np.float32]
    for dt in dts:
        run(dt, only)
