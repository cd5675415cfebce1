"""GPU: run_oLBFGS with HOST pointers above 32 MiB per vector - the PCIe copies of grad / x are cut into pieces that
overlap K1 / K3 / K4 (stochqn_b200.cu: take_step_host / pair_host).  Same task sequence and iterates as the oracle
(stochqn.c:978-1036), with a ragged last piece, with the default options (device mirror of x trusted, no write-back of
the direction into the host `grad`) and with the reference's literal behaviour restored by the options."""
import numpy as np
import pytest

from cuda_stepper import CudaStepper
from oracle import stochqn_np as O
from oracle.driver import HostStepper, discrete, run_trace
from oracle.problems import Rosenbrock
from stochqn_b200 import _lib

pytestmark = pytest.mark.gpu
N = (1 << 22) + 5          # 33.5 MB per fp64 vector: above the pipeline threshold, not a multiple of anything
KW = dict(mem_size=3, hess_init=0.0, y_reg=0.0, min_curvature=1e-4, check_nan=1)


def _oracle(calls, hooks=None):
    p = Rosenbrock(N)
    return run_trace(HostStepper(O.OracleOLBFGS(N, **KW), p.x0()), p, calls, 1e-4, hooks=hooks, keep_x=True)


def _err(ta, tb, key="x"):
    return max(float(np.max(np.abs(a[key] - b[key])) / max(np.max(np.abs(b[key])), 1e-300)) for a, b in zip(ta, tb))


@pytest.mark.parametrize("chunk_mb", ["8", "64"])
def test_pipelined_host_calls_match_the_oracle(monkeypatch, chunk_mb):
    monkeypatch.setenv("STOCHQN_B200_STAGE_CHUNK_MB", chunk_mb)
    to = _oracle(15)
    p = Rosenbrock(N)
    sc = CudaStepper("oLBFGS", p.x0(), mode="host", **KW)
    tc = run_trace(sc, p, 15, 1e-4, keep_x=True)
    assert discrete(to) == discrete(tc)
    assert _err(tc, to) <= 1e-10
    assert sc.abi.lib.stochqn_b200_get_option(sc.ws, _lib.OPT_GRAD_WRITEBACK) == -1
    assert sc.abi.lib.stochqn_b200_get_option(sc.ws, _lib.OPT_TRUST_X_MIRROR) == 1
    sc.close()


def test_pipelined_host_calls_with_the_literal_reference_behaviour(monkeypatch):
    """write-back on: `grad` holds -step*direction after a step (stochqn.c:1006); mirror not trusted: a caller that edits x
    between calls is honoured."""
    monkeypatch.setenv("STOCHQN_B200_STAGE_CHUNK_MB", "8")

    def nudge(stepper, task, payload):           # the caller modifies x between two calls
        stepper.x[7] += 1e-3

    to = _oracle(13, hooks={9: nudge})
    p = Rosenbrock(N)
    sc = CudaStepper("oLBFGS", p.x0(), mode="host", grad_writeback=1, **KW)
    sc.abi.lib.stochqn_b200_set_option(sc.ws, _lib.OPT_TRUST_X_MIRROR, 0)
    tc = run_trace(sc, p, 13, 1e-4, hooks={9: nudge}, keep_x=True)
    assert discrete(to) == discrete(tc)
    assert _err(tc, to) <= 1e-10
    steps = [(a, b) for a, b in zip(tc, to) if b["ret"] == 1]
    assert len(steps) >= 5 and _err([a for a, _ in steps], [b for _, b in steps], key="grad") <= 1e-9
    sc.close()


def test_nan_gradient_in_the_pipelined_path_is_rejected():
    def bad(stepper, task, payload):
        payload["grad"] = payload["grad"].copy()
        payload["grad"][N - 2] = np.nan          # in the ragged last piece

    to = _oracle(11, hooks={7: bad})
    p = Rosenbrock(N)
    sc = CudaStepper("oLBFGS", p.x0(), mode="host", **KW)
    tc = run_trace(sc, p, 11, 1e-4, hooks={7: bad}, keep_x=True)
    assert discrete(to) == discrete(tc)
    assert tc[7]["info"] == 203 and tc[7]["mem_used"] == 0
    assert _err(tc, to) <= 1e-10
    sc.close()
