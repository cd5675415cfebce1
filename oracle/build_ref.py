"""TEST INFRASTRUCTURE ONLY - builds the CPU oracle; never imported by the product.

Compiles the UNMODIFIED reference core where it lies (/root/reference/src/stochqn.c,
/root/reference/include/stochqn.h) into ``oracle/_ref/``:

    libstochqn_ref_f64.so / libstochqn_ref_f32.so   the reference C library (USE_DOUBLE / USE_FLOAT)
    rosen_harness_f64                               oracle/rosen_harness.c linked to it (CPU baseline)

No reference source is copied into the repository and the reference's own build
system (CMake / setup.py / R CMD) is not run.  Two pre-include shims from
``oracle/shim/`` stand in for generated / host-language pieces:

    blasfuns.h      the header stochqn.c:79 includes; maps cblas_* onto SciPy's bundled
                    OpenBLAS (the only BLAS in the image)
    zero_malloc.h   malloc -> calloc, i.e. the R allocators' zero-filled backup buffers
                    (R/allocators.R:8-9), which makes quirk Q1 deterministic

Flags follow the reference's CMakeLists.txt:86 / setup.py:21 (-O2 -fopenmp -std=c99)
except that ``-march=native`` is replaced by ``-march=x86-64-v3``: the artefacts are
built in the CPU container and travel to the GPU box, whose host CPU may differ.
OpenBLAS picks its kernels at run time, so BLAS speed is unaffected.

``oracle/_ref/`` is git-ignored (never in history) but not gpurun-ignored.
/root/reference does not exist on the GPU box: there the prebuilt files are used as is.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("STOCHQN_REFERENCE", "/root/reference")


def openblas_path() -> str:
    import scipy  # noqa: F401  (only to locate the wheel's bundled library)

    libs = os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs")
    found = sorted(glob.glob(os.path.join(libs, "libscipy_openblas*.so")))
    if not found:
        raise RuntimeError("SciPy's bundled OpenBLAS not found under " + libs)
    return found[0]


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF, "src", "stochqn.c"))


def artefacts() -> dict:
    return {
        "f64": os.path.join(OUT, "libstochqn_ref_f64.so"),
        "f32": os.path.join(OUT, "libstochqn_ref_f32.so"),
        "harness_f64": os.path.join(OUT, "rosen_harness_f64"),
    }


def _run(cmd):
    subprocess.run(cmd, check=True)


def build(force: bool = False, verbose: bool = False) -> dict:
    """Build (or reuse) the reference artefacts.  Returns {name: path} of what exists."""
    art = artefacts()
    if not reference_available():
        return {k: v for k, v in art.items() if os.path.exists(v)}
    os.makedirs(OUT, exist_ok=True)
    ob = openblas_path()
    common = ["gcc", "-O2", "-fopenmp", "-march=x86-64-v3", "-std=gnu99", "-fPIC",
              "-I" + os.path.join(HERE, "shim")]
    src = os.path.join(REF, "src", "stochqn.c")
    for tag, macro in (("f64", "-DUSE_DOUBLE"), ("f32", "-DUSE_FLOAT")):
        out = art[tag]
        if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
            cmd = common + ["-shared", macro, "-include", os.path.join(HERE, "shim", "zero_malloc.h"),
                            "-I" + os.path.join(REF, "include"), src, "-o", out, ob,
                            "-Wl,-rpath," + os.path.dirname(ob), "-lm"]
            if verbose:
                print(" ".join(cmd))
            _run(cmd)
    hsrc = os.path.join(HERE, "rosen_harness.c")
    hout = art["harness_f64"]
    if force or not os.path.exists(hout) or os.path.getmtime(hout) < os.path.getmtime(hsrc):
        # the harness only needs struct layouts / prototypes: our own header is
        # source-compatible with the reference's, so it compiles against either
        cmd = ["gcc", "-O2", "-fopenmp", "-march=x86-64-v3", "-std=gnu99", "-DUSE_DOUBLE",
               "-I" + os.path.join(REPO, "include"), hsrc, "-o", hout,
               art["f64"], ob, "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + os.path.dirname(ob), "-lm"]
        if verbose:
            print(" ".join(cmd))
        _run(cmd)
    return {k: v for k, v in art.items() if os.path.exists(v)}


if __name__ == "__main__":
    got = build(force="--force" in sys.argv, verbose=True)
    for k, v in got.items():
        print(k, v)
    if not got:
        sys.exit("reference sources not found and no prebuilt oracle/_ref present")
