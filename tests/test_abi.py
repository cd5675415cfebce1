"""CPU: the drop-in boundary.  The CUDA libraries load without a GPU and export every symbol the two
public headers declare; the struct layouts of include/stochqn.h equal the ctypes mirrors (and the
reference's own header where it is available).  No compute entry point is called here."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from stochqn_b200 import _abi, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")


def _declared_functions(header):
    txt = open(os.path.join(INC, header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    txt = txt.split("#ifdef __cplusplus\n#include <new>")[0]          # C part only
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}]*\)\s*;", txt)
    return sorted(set(n for n in names if not n.startswith("defined")))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_library_loads_and_exports_every_declared_symbol(dtype):
    abi = _lib.load(dtype)
    declared = _declared_functions("stochqn.h") + _declared_functions("stochqn_b200.h")
    assert "run_oLBFGS" in declared and "stochqn_b200_set_comm" in declared and len(declared) >= 30
    for name in declared:
        assert hasattr(abi.lib, name), "library does not export %s" % name
    assert abi.lib.stochqn_b200_real_bytes() == np.dtype(dtype).itemsize
    assert set(_abi.StochqnABI.REFERENCE_SYMBOLS) <= set(declared)


def _layout_from_header(include_dir, macro):
    src = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "stochqn.h"
    #define P(T, f) printf(#T "." #f " %zu\n", offsetof(T, f));
    int main(void) {
        printf("sizeof.bfgs_mem %zu\n", sizeof(bfgs_mem)); printf("sizeof.fisher_mem %zu\n", sizeof(fisher_mem));
        printf("sizeof.workspace_oLBFGS %zu\n", sizeof(workspace_oLBFGS)); printf("sizeof.workspace_SQN %zu\n", sizeof(workspace_SQN));
        printf("sizeof.workspace_adaQN %zu\n", sizeof(workspace_adaQN));
        P(bfgs_mem, s_mem) P(bfgs_mem, y_mem) P(bfgs_mem, s_bak) P(bfgs_mem, mem_size) P(bfgs_mem, mem_used) P(bfgs_mem, mem_st_ix)
        P(bfgs_mem, upd_freq) P(bfgs_mem, y_reg) P(bfgs_mem, min_curvature)
        P(fisher_mem, F) P(fisher_mem, mem_size) P(fisher_mem, mem_used) P(fisher_mem, mem_st_ix)
        P(workspace_oLBFGS, bfgs_memory) P(workspace_oLBFGS, grad_prev) P(workspace_oLBFGS, hess_init) P(workspace_oLBFGS, niter)
        P(workspace_oLBFGS, section) P(workspace_oLBFGS, check_nan) P(workspace_oLBFGS, n)
        P(workspace_SQN, x_sum) P(workspace_SQN, x_avg_prev) P(workspace_SQN, use_grad_diff) P(workspace_SQN, niter) P(workspace_SQN, section) P(workspace_SQN, n)
        P(workspace_adaQN, fisher_memory) P(workspace_adaQN, H0) P(workspace_adaQN, grad_sum_sq) P(workspace_adaQN, f_prev) P(workspace_adaQN, max_incr)
        P(workspace_adaQN, scal_reg) P(workspace_adaQN, rmsprop_weight) P(workspace_adaQN, use_grad_diff) P(workspace_adaQN, niter) P(workspace_adaQN, section) P(workspace_adaQN, n)
        printf("enum.calc_grad %d\n", (int) calc_grad); printf("enum.calc_hess_vec %d\n", (int) calc_hess_vec);
        printf("enum.invalid_input %d\n", (int) invalid_input); printf("enum.func_increased %d\n", (int) func_increased);
        printf("enum.search_direction_was_nan %d\n", (int) search_direction_was_nan); printf("enum.received_invalid_input %d\n", (int) received_invalid_input);
        return 0; }
    '''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.run(["gcc", "-std=c99", macro, "-I" + include_dir, c, "-o", exe], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    return dict(line.split() for line in out.strip().splitlines())


@pytest.mark.parametrize("macro,real", [("-DUSE_DOUBLE", C.c_double), ("-DUSE_FLOAT", C.c_float)])
def test_struct_layout_matches_ctypes_and_reference_header(macro, real):
    mine = _layout_from_header(INC, macro)
    S = _abi.make_structs(real)
    for key, val in mine.items():
        kind, field = key.split(".")
        if kind == "sizeof":
            assert C.sizeof(S[field]) == int(val), key
        elif kind != "enum":
            assert getattr(S[kind], field).offset == int(val), key
    assert mine["enum.calc_grad"] == "101" and mine["enum.calc_hess_vec"] == "104" and mine["enum.invalid_input"] == "100"
    assert mine["enum.func_increased"] == "201" and mine["enum.search_direction_was_nan"] == "203"
    assert mine["enum.received_invalid_input"] == "-1000"
    ref_inc = "/root/reference/include"
    if os.path.exists(os.path.join(ref_inc, "stochqn.h")):
        assert _layout_from_header(ref_inc, macro) == mine       # byte-identical layout to the reference's header


def test_cpp_classes_compile_against_our_header():
    """The RAII classes of the reference header (oLBFGS / SQN / adaQN, include/stochqn.h:400-508) exist with the
    same constructor defaults and methods: a C++ translation unit written against the reference compiles."""
    src = r'''
    #include "stochqn.h"
    int use(double* x, double* g, double* hv) {
        oLBFGS a(10); SQN b(10, 5, 3, 0.0, 0, 1e-8, 1, 1); adaQN c(10);
        a.run(1e-3, x, g); b.run(1e-3, x, g, hv); c.run(1e-3, x, 0.0, g);
        return (int) a.get_task() + (int) b.get_iter_info() + (int) c.get_n_iter() + (b.get_req_vec() != 0) + (a.get_req() != 0);
    }'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.cpp")
        open(c, "w").write(src)
        subprocess.run(["g++", "-std=c++11", "-c", "-I" + INC, c, "-o", os.path.join(d, "t.o")], check=True)


def test_no_cpu_fallback_without_a_device():
    """On a box without CUDA the constructors fail loudly (NULL + message), they do not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    abi = _lib.load(np.float64)
    ws = abi.lib.initialize_oLBFGS(16, 3, 0.0, 0.0, 0.0, 1, 1)
    assert not ws
    assert "CUDA" in _lib.last_error(abi) or "device" in _lib.last_error(abi)


@pytest.mark.parametrize("macro,dtype", [("-DUSE_DOUBLE", np.float64), ("-DUSE_FLOAT", np.float32)])
def test_extension_struct_layouts_match_ctypes(macro, dtype):
    """stochqn_b200_rows / stochqn_b200_model / stochqn_b200_fit_report / stochqn_b200_host_state (include/stochqn_b200.h)
    against the ctypes mirrors the Python layer passes to stochqn_b200_fit_batch / export / import."""
    src = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "stochqn_b200.h"
    #define P(T, f) printf(#T "." #f " %zu\n", offsetof(T, f));
    int main(void) {
        printf("sizeof.stochqn_b200_rows %zu\n", sizeof(stochqn_b200_rows));
        printf("sizeof.stochqn_b200_model %zu\n", sizeof(stochqn_b200_model));
        printf("sizeof.stochqn_b200_fit_report %zu\n", sizeof(stochqn_b200_fit_report));
        printf("sizeof.stochqn_b200_host_state %zu\n", sizeof(stochqn_b200_host_state));
        P(stochqn_b200_rows, X) P(stochqn_b200_rows, ldx) P(stochqn_b200_rows, y) P(stochqn_b200_rows, ldy) P(stochqn_b200_rows, sw) P(stochqn_b200_rows, nrows)
        P(stochqn_b200_model, model) P(stochqn_b200_model, fit_intercept) P(stochqn_b200_model, ncols) P(stochqn_b200_model, nclasses)
        P(stochqn_b200_model, reg_param) P(stochqn_b200_model, work)
        P(stochqn_b200_fit_report, calls) P(stochqn_b200_fit_report, n_info) P(stochqn_b200_fit_report, last_info)
        P(stochqn_b200_fit_report, x_changed) P(stochqn_b200_fit_report, long_batch_used)
        P(stochqn_b200_host_state, s_mem) P(stochqn_b200_host_state, F)
        return 0; }
    '''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.run(["gcc", "-std=c99", macro, "-I" + INC, c, "-o", exe], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    lay = dict(line.split() for line in out.strip().splitlines())
    abi = _lib.load(dtype)
    mirrors = {"stochqn_b200_rows": _lib.Rows, "stochqn_b200_model": abi.Model, "stochqn_b200_fit_report": _lib.FitReport,
               "stochqn_b200_host_state": _lib.HostState}
    for key, val in lay.items():
        kind, field = key.split(".")
        if kind == "sizeof":
            assert C.sizeof(mirrors[field]) == int(val), key
        else:
            assert getattr(mirrors[kind], field).offset == int(val), key


def test_fit_batch_refuses_a_foreign_workspace():
    """stochqn_b200_fit_batch answers a pointer it did not create with -1000 (the reference's invalid-input code),
    without touching a GPU."""
    abi = _lib.load(np.float64)
    fake = (C.c_char * 256)()
    task, req, req_vec = C.c_int(101), C.c_void_p(), C.c_void_p()
    rows = _lib.Rows()
    M = abi.Model()
    rep = _lib.FitReport()
    rc = abi.lib.stochqn_b200_fit_batch(C.cast(fake, C.c_void_p), C.cast(fake, C.c_void_p), 0.1, C.byref(M), C.byref(rows), None, None,
                                        C.byref(task), C.byref(req), C.byref(req_vec), C.byref(rep))
    assert rc == -1000


def test_multinomial_work_size_is_monotone_in_the_batch_size():
    """A work buffer sized for the largest batch (the long batch of a guided fit) must hold the plan of every smaller batch: the
    size function may not decrease when rows are added (the number of split-k partial products does)."""
    import numpy as np
    from stochqn_b200 import _lib
    for dtype in (np.float64, np.float32):
        lib = _lib.load(dtype).lib
        for d, K in ((1836, 159), (8192, 4096), (40, 7), (300, 512)):
            prev = 0
            for B in list(range(1, 300)) + list(range(300, 20000, 97)):
                cur = lib.stochqn_b200_multinomial_work_size(B, d, K)
                assert cur >= prev, (dtype, d, K, B, prev, cur)
                prev = cur
