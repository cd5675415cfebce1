"""GPU: the two arms of bench.py on the same workload - the reference C library driven by oracle/rosen_harness.c and
the CUDA library driven by bench.CudaRosen - end with the same iterate: n = 2^20, mem_size 10, 12 warm-up + 8 more
iterations; counters identical, norm / sum / probe entries within 1e-10 (reference: stochqn.c:978-1036)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "rosen_harness_f64")), reason="oracle/_ref not built")
@pytest.mark.parametrize("n", [1 << 20, (1 << 18) + 3])
def test_both_arms_reach_the_same_iterate(n):
    import torch

    import bench
    from stochqn_b200 import _lib

    ref = bench._run_harness(n, 12, 8, os.cpu_count() or 1)
    run = bench.CudaRosen(torch, _lib.load(np.float64), n)
    run.run(20)
    mine = run.probes()
    run.close()
    assert mine["probe_idx"] == ref["probe_idx"]
    for k in ("info_events", "mem_used", "mem_st_ix", "niter"):
        assert mine[k] == ref[k], (k, mine[k], ref[k])
    assert bench.rel_probe_distance(mine, ref) <= 1e-10


def test_config_is_the_same_from_both_arms():
    import bench
    assert bench.config_dict(1 << 27, 4) == bench.config_dict(1 << 27, 4)
    assert bench.effective_warmup(5) == 12 and bench.effective_warmup(30) == 30
