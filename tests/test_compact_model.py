"""CPU: the compact-form algebra + Gram bookkeeping the kernels use equals the reference's
two-loop recursion (stochqn.c:663-708) for every ring-buffer state, including partially filled
memory, wrap-around, scalar / last-pair / diagonal H0 and the rejection quirk Q1."""
import numpy as np
import pytest

from compact_model import CompactModel
from oracle import stochqn_np as O


def _fill(rng, n, m, dtype=np.float64):
    mem = O.BfgsMem(m, n, 0.0, 0.0, 1, dtype)
    return mem


@pytest.mark.parametrize("h0", [0.0, 0.37])
@pytest.mark.parametrize("m", [1, 3, 10])
def test_scalar_h0_matches_two_loop_through_ring_states(m, h0):
    rng = np.random.default_rng(5 + m)
    n = 200
    mem = _fill(rng, n, m)
    cm = CompactModel(m)
    A = rng.standard_normal((n, n)); A = A @ A.T / n + np.eye(n)
    for it in range(3 * m + 2):
        g = rng.standard_normal(n)
        d_model, U = cm.direction(g, mem.s_mem, mem.y_mem, mem.mem_used, mem.mem_st_ix, h0)
        d_ref = g.copy()
        if mem.mem_used > 0:
            oldest = 0 if mem.mem_st_ix == mem.mem_used else mem.mem_st_ix
            O.approx_inv_hess_grad(d_ref, None, h0, mem, oldest)
        assert np.max(np.abs(d_model - d_ref)) <= 1e-11 * np.max(np.abs(d_ref))
        assert U >= np.linalg.norm(d_ref) * (1 - 1e-12)
        # add a pair with positive curvature
        s = rng.standard_normal(n) * 0.1
        slot = mem.mem_st_ix
        mem.s_mem[slot] = s
        mem.y_mem[slot] = A @ s
        O.incr_bfgs_counters(mem)
        cm.pending = slot


def test_diag_h0_matches_two_loop():
    rng = np.random.default_rng(11)
    n, m = 150, 4
    mem = _fill(rng, n, m)
    cm = CompactModel(m)
    A = rng.standard_normal((n, n)); A = A @ A.T / n + np.eye(n)
    for it in range(11):
        g = rng.standard_normal(n)
        h = g / np.sqrt(rng.random(n) + 1e-4)          # sign-indefinite, like quirk Q2
        d_model, U = cm.direction_diag(g, h, mem.s_mem, mem.y_mem, mem.mem_used, mem.mem_st_ix)
        if mem.mem_used > 0:
            d_ref = g.copy()
            oldest = 0 if mem.mem_st_ix == mem.mem_used else mem.mem_st_ix
            O.approx_inv_hess_grad(d_ref, h, 0.0, mem, oldest)
        else:
            d_ref = h.copy()
        assert np.max(np.abs(d_model - d_ref)) <= 1e-10 * np.max(np.abs(d_ref))
        assert U >= np.linalg.norm(d_ref) * (1 - 1e-12)
        s = rng.standard_normal(n) * 0.1
        slot = mem.mem_st_ix
        mem.s_mem[slot] = s
        mem.y_mem[slot] = A @ s
        O.incr_bfgs_counters(mem)
        cm.pending = slot


def test_zeroed_oldest_slot_gives_nonfinite_direction():
    """Quirk Q1: a rejected pair with full memory leaves s = y = 0 in the oldest slot; the next
    direction must come out non-finite in both formulations (-> search_direction_was_nan)."""
    rng = np.random.default_rng(2)
    n, m = 50, 2
    mem = _fill(rng, n, m)
    cm = CompactModel(m)
    for _ in range(2):
        s = rng.standard_normal(n); slot = mem.mem_st_ix
        mem.s_mem[slot] = s; mem.y_mem[slot] = 2.0 * s
        O.incr_bfgs_counters(mem); cm.pending = slot
        cm.fold(mem.s_mem, mem.y_mem, mem.mem_used)
    slot = mem.mem_st_ix            # full memory: this is the oldest pair
    mem.s_mem[slot] = 0; mem.y_mem[slot] = 0
    cm.zero_slot(slot)
    g = rng.standard_normal(n)
    d_model, U = cm.direction(g, mem.s_mem, mem.y_mem, mem.mem_used, mem.mem_st_ix, 0.0)
    d_ref = g.copy()
    O.approx_inv_hess_grad(d_ref, None, 0.0, mem, mem.mem_st_ix)
    assert not np.all(np.isfinite(d_ref))
    assert not np.all(np.isfinite(d_model)) and not np.isfinite(U)
