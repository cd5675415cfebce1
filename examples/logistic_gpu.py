"""Example: logistic regression with the stochastic quasi-Newton optimizers on a GPU-resident model matrix.

The same calls as the reference's README (``StochasticLogisticRegression(...).fit(X, y)``, guided ``SQN(...).fit``
with user callbacks, free-mode ``oLBFGS_free().run_optimizer``), with the data on the device.  Needs a CUDA device:

    python examples/logistic_gpu.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stochqn_b200.guided import SQN                      # noqa: E402
from stochqn_b200.logistic import StochasticLogisticRegression, grad_fun_bin, hessvec_fun_bin, obj_fun_bin  # noqa: E402
from stochqn_b200.optimizers import oLBFGS_free           # noqa: E402


def main():
    torch.manual_seed(0)
    n, d = 200_000, 512
    X = torch.randn(n, d, device="cuda", dtype=torch.float64) / d ** 0.5
    w_true = torch.randn(d, device="cuda", dtype=torch.float64) * 3
    y = torch.where(torch.rand(n, device="cuda", dtype=torch.float64) < torch.sigmoid(X @ w_true), 1.0, -1.0)

    # 1. the estimator: every callback is a kernel of the library, the request loop of a mini-batch runs inside it
    m = StochasticLogisticRegression(optimizer="SQN", step_size=1e-1, reg_param=1e-4, batches_per_epoch=50, nepochs=3,
                                     bfgs_upd_freq=10, valset_frac=0.1, verbose=False)
    m.fit(X, y)
    acc = ((m.predict(X) > 0) == (y > 0)).double().mean().item()
    print("StochasticLogisticRegression: %d iterations, train accuracy %.3f" % (m.optimizer.niter, acc))

    # 2. guided mode with user callbacks (here: the bundled device callbacks passed explicitly)
    sw = torch.full((n,), 1.0 / n, device="cuda", dtype=torch.float64)
    opt = SQN(torch.zeros(d + 1, device="cuda", dtype=torch.float64), grad_fun_bin, obj_fun=obj_fun_bin,
              hess_vec_fun=hessvec_fun_bin, batches_per_epoch=50, nepochs=2, step_size=1e-1, decr_step_size=None,
              bfgs_upd_freq=10, verbose=False)
    opt.fit(X, y, sw, additional_kwargs={"reg_param": 1e-4})
    print("guided SQN: %d iterations, objective %.5f" % (opt.niter, obj_fun_bin(opt.x, X, y, sample_weight=sw, reg_param=1e-4)))

    # 3. free mode: the caller serves the requests (the reference's request loop, stochqn/_optimizers.py:988-1045)
    free = oLBFGS_free(mem_size=10)
    x = torch.zeros(d + 1, device="cuda", dtype=torch.float64)
    req = free.run_optimizer(x, 1e-1)
    batch = 0
    while free.niter < 100:
        if req["task"] == "calc_grad":
            batch = (batch + 1) % 50
        rows = slice(batch * 4000, (batch + 1) * 4000)
        free.update_gradient(grad_fun_bin(req["requested_on"], X[rows], y[rows], sample_weight=sw[rows] * 50, reg_param=1e-4))
        req = free.run_optimizer(x, 1e-1)
    print("free-mode oLBFGS: %d iterations, |x| = %.4f" % (free.niter, float(torch.linalg.vector_norm(x))))


if __name__ == "__main__":
    main()
