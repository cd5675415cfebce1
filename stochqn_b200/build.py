"""Build the CUDA libraries in-tree (``stochqn_b200/lib/``) with nvcc for sm_100a.

    libstochqn_b200_f64.so   -DUSE_DOUBLE   (real_t = double, the reference's default)
    libstochqn_b200_f32.so   -DUSE_FLOAT

Both export the reference's C ABI (include/stochqn.h) plus the extensions of
include/stochqn_b200.h.  nvcc cross-compiles without a GPU; the .so files are git-ignored
but travel to the GPU box with the snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")

SOURCES = ["stochqn_b200.cu", "callbacks.cu", "multinomial.cu"]
DEPS = ["kernels.cuh", "kernels_small.cuh", "kernels_adaqn.cuh", "kernels_loop.cuh", "kernels_fit.cuh", "mn_small.cuh", "logistic_form.cuh", "p2p.cuh", "gemm_tf32_sm100.cuh", "vecio.cuh", "adaqn_impl.inc", "ext_impl.inc", "guided_impl.inc", "internal_rs.h",
        os.path.join(REPO, "include", "stochqn.h"), os.path.join(REPO, "include", "stochqn_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-I" + os.path.join(REPO, "include"), "-I" + CSRC]


def lib_path(tag: str) -> str:
    return os.path.join(LIBDIR, "libstochqn_b200_%s.so" % tag)


def _newest_dep() -> float:
    t = 0.0
    for d in SOURCES + DEPS:
        p = d if os.path.isabs(d) else os.path.join(CSRC, d)
        t = max(t, os.path.getmtime(p))
    return max(t, os.path.getmtime(__file__))


def _compile(args):
    tag, macro, src, verbose, extra = args
    obj = os.path.join(OBJDIR, "%s_%s.o" % (os.path.splitext(src)[0], tag))
    cmd = ["nvcc"] + NVCC_FLAGS + extra + [macro, "-c", os.path.join(CSRC, src), "-o", obj]
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s (%s):\n%s\n%s" % (src, tag, r.stdout, r.stderr))
    return obj, r.stderr


def build(force: bool = False, verbose: bool = False, ptxas_info: bool = False) -> dict:
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    newest = _newest_dep()
    tags = {"f64": "-DUSE_DOUBLE", "f32": "-DUSE_FLOAT"}
    todo = [t for t in tags if force or not os.path.exists(lib_path(t)) or os.path.getmtime(lib_path(t)) < newest]
    extra = ["-Xptxas", "-v"] if ptxas_info else []
    jobs = [(t, tags[t], s, verbose, extra) for t in todo for s in SOURCES]
    logs = {}
    if jobs:
        with ThreadPoolExecutor(max_workers=min(6, len(jobs))) as ex:
            results = list(ex.map(_compile, jobs))
        for (t, _, s, _, _), (obj, log) in zip(jobs, results):
            logs[(t, s)] = log
        for t in todo:
            objs = [os.path.join(OBJDIR, "%s_%s.o" % (os.path.splitext(s)[0], t)) for s in SOURCES]
            cmd = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib_path(t)] + objs + ["-ldl"]
            if verbose:
                print(" ".join(cmd), flush=True)
            subprocess.run(cmd, check=True)
    out = {t: lib_path(t) for t in tags}
    out["_ptxas"] = logs
    # host-side helper for the end-to-end bench leg (CPU gradient callback, plain C + OpenMP)
    hsrc = os.path.join(REPO, "tools", "host_callbacks.c")
    hout = os.path.join(LIBDIR, "libhostcb_f64.so")
    if force or not os.path.exists(hout) or os.path.getmtime(hout) < os.path.getmtime(hsrc):
        subprocess.run(["gcc", "-O3", "-fopenmp", "-march=x86-64-v3", "-ffp-contract=off", "-fPIC", "-shared", hsrc, "-o", hout], check=True)
    out["hostcb_f64"] = hout
    return out


if __name__ == "__main__":
    res = build(force="--force" in sys.argv, verbose=True, ptxas_info="--ptxas" in sys.argv)
    if "--ptxas" in sys.argv:
        for k, v in res["_ptxas"].items():
            print("=====", k)
            print(v)
    for k, v in res.items():
        if not k.startswith("_"):
            print(k, v)
