"""CPU: the Cython binding (bindings/cython/stochqn_b200_cy.pyx) cythonizes and compiles against the public headers and
links against the double library; the module imports without a GPU and fails loudly (MemoryError) on construction."""
import os
import subprocess
import sys
import sysconfig
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PYX = os.path.join(ROOT, "bindings", "cython", "stochqn_b200_cy.pyx")

cython = pytest.importorskip("Cython")


def test_cython_binding_builds_and_imports():
    from stochqn_b200 import _lib
    libdir = os.path.dirname(_lib.lib_path(np.float64))
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "stochqn_b200_cy.c")
        subprocess.run([sys.executable, "-m", "cython", "-3", PYX, "-o", c], check=True)
        so = os.path.join(d, "stochqn_b200_cy" + sysconfig.get_config_var("EXT_SUFFIX"))
        cmd = ["gcc", "-O1", "-shared", "-fPIC", "-DUSE_DOUBLE", "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION",
               "-I" + sysconfig.get_paths()["include"], "-I" + np.get_include(), "-I" + os.path.join(ROOT, "include"),
               c, "-o", so, "-L" + libdir, "-lstochqn_b200_f64", "-Wl,-rpath," + libdir]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
        code = ("import sys; sys.path.insert(0, %r); import stochqn_b200_cy as m; import torch\n"
                "ok = all(hasattr(m, n) for n in ('py_init_oLBFGS','py_init_SQN','py_init_adaQN','py_run_oLBFGS','py_run_SQN','py_run_adaQN','py_run_oLBFGS_ptr'))\n"
                "assert ok\n"
                "if not torch.cuda.is_available():\n"
                "    try:\n"
                "        m.py_init_oLBFGS(8, 3, 0.0, 0.0, 0.0, 1, 1); raise SystemExit(3)\n"
                "    except MemoryError:\n"
                "        pass\n"
                "print('cython binding ok')\n") % d
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
        assert r.returncode == 0 and "cython binding ok" in r.stdout, r.stdout + r.stderr[-2000:]
