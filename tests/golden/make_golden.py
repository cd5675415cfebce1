"""Generate tests/golden/traces_f64.json from the REFERENCE C library itself (oracle/_ref, built from
/root/reference/src/stochqn.c by oracle/build_ref.py).  Run in the build container (the reference
sources are not available on the GPU box); the output is committed.

For every case of tests/cases.py: the per-call discrete trace (task, ret, info, req label, niter,
section, mem_used, mem_st_ix, Fisher counters), ||x|| after every call and the final x.
Plus the known-answer run of the reference's example program (example/c_rosen.c: SQN, n = 4).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

from cases import CASES  # noqa: E402
from oracle import ref_lib as R  # noqa: E402
from oracle.driver import DISCRETE_FIELDS, HostStepper, run_trace  # noqa: E402
from oracle.problems import Rosenbrock  # noqa: E402

REF = {"oLBFGS": R.RefOLBFGS, "SQN": R.RefSQN, "adaQN": R.RefAdaQN}


def c_rosen_example():
    """example/c_rosen.c:71-126 replayed through the reference library: SQN, n=4, mem_size 5, L=3,
    min_curvature 0, y_reg 1e-8, step 1e-3, until niter reaches 200."""
    prob = Rosenbrock(4, example_quirk=True)
    x = np.array([1.3, 0.7, 0.8, 1.9])
    opt = R.RefSQN(4, 5, 3, 0.0, 0, 1e-8, 1, 1)
    st = HostStepper(opt, x)
    f0 = prob.fun(st.x)
    ret, task, info = st.call(1e-3)
    printed = {}
    while opt.niter < 200:
        if task == 101:
            st.write("grad", prob.grad(st.read("req"), "new"))
        elif task == 104:
            st.write("hess_vec", prob.hess_vec(st.read("req"), st.read("req_vec")))
        ret, task, info = st.call(1e-3)
        if ret and (opt.niter + 1) % 10 == 0:
            printed[str(opt.niter + 1)] = prob.fun(st.x)
    return {"f_initial": f0, "f_printed": printed, "f_final": prob.fun(st.x), "x_final": st.x.tolist()}


def main():
    out = {"fields": list(DISCRETE_FIELDS), "cases": {}}
    for name, kind, kw, prob_f, calls, step in CASES:
        p = prob_f()
        st = HostStepper(REF[kind](len(p.x0()), dtype=np.float64, **kw), p.x0())
        tr = run_trace(st, p, calls, step, keep_x=True)
        out["cases"][name] = {
            "discrete": [[r.get(k) for k in DISCRETE_FIELDS] for r in tr],
            "x_norm": [r["x_norm"] for r in tr],
            "x_final": tr[-1]["x"].tolist(),
        }
    out["c_rosen_example"] = c_rosen_example()
    with open(os.path.join(HERE, "traces_f64.json"), "w") as f:
        json.dump(out, f)
    print("wrote", os.path.join(HERE, "traces_f64.json"), os.path.getsize(os.path.join(HERE, "traces_f64.json")), "bytes")
    ex = out["c_rosen_example"]
    print("c_rosen: f0=%.4f f10=%.4f f200=%.4f final=%.4f x=%s" % (ex["f_initial"], ex["f_printed"]["10"], ex["f_printed"]["200"],
                                                                  ex["f_final"], np.round(ex["x_final"], 6)))


if __name__ == "__main__":
    main()
