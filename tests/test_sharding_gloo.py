"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path.

* shard_bounds tiles [0, n) contiguously for any (n, world);
* the 128-byte communicator id travels from rank 0 to the others (broadcast_id);
* a step computed from per-shard partial sums that are all-reduced (what K1 + one NCCL all-reduce + K2 do
  on the GPUs) equals the unsharded oracle step - including the Rosenbrock halo exchange.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stochqn_b200.distributed import broadcast_id, shard_bounds


def test_shard_bounds_tile_the_vector():
    for n in (1, 7, 1000, 1001, 2 ** 20 + 3):
        for world in (1, 2, 3, 8):
            pos = 0
            for r in range(world):
                off, cnt = shard_bounds(n, r, world)
                assert off == pos and cnt >= 0
                pos += cnt
            assert pos == n
            sizes = [shard_bounds(n, r, world)[1] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, m, q):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    sys.path.insert(0, os.path.dirname(here))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.problems import Rosenbrock
        ident = broadcast_id(bytes(range(128)) if rank == 0 else b"", rank)
        assert ident == bytes(range(128))

        off, cnt = shard_bounds(n, rank, world)
        rng = np.random.default_rng(0)                # same stream on every rank: identical global data
        S = rng.standard_normal((m, n)) * 0.1
        A = np.linspace(1.0, 3.0, n)
        Y = S * A                                    # positive curvature pairs
        x = Rosenbrock(n).x0()
        sl = slice(off, off + cnt)

        # --- halo exchange: every rank contributes (first, last) to a zero-padded record, sum = all-gather
        rec = torch.zeros(2 * world, dtype=torch.float64)
        rec[2 * rank], rec[2 * rank + 1] = x[off], x[off + cnt - 1]
        dist.all_reduce(rec)
        left = rec[2 * (rank - 1) + 1].item() if rank > 0 else 0.0
        right = rec[2 * (rank + 1)].item() if rank < world - 1 else 0.0
        xe = np.concatenate(([left], x[sl], [right]))
        g_loc = np.zeros(cnt)
        for i in range(cnt):
            gi = off + i
            if gi > 0:
                g_loc[i] += 200.0 * (xe[i + 1] - xe[i] ** 2)
            if gi < n - 1:
                g_loc[i] -= 400.0 * (xe[i + 2] - xe[i + 1] ** 2) * xe[i + 1] + 2.0 * (1.0 - xe[i + 1])
        g_full = Rosenbrock(n).grad(x)
        assert np.allclose(g_loc, g_full[sl], rtol=1e-13, atol=1e-13)

        # --- K1 on the shard: the sum record [S'g, Y'g, full Gram columns here, g'g], then ONE all-reduce
        part = np.concatenate([S[:, sl] @ g_loc, Y[:, sl] @ g_loc, (S[:, sl] @ Y[:, sl].T).ravel(),
                               (Y[:, sl] @ Y[:, sl].T).ravel(), [g_loc @ g_loc]])
        t = torch.from_numpy(part.copy())
        dist.all_reduce(t)
        tot = t.numpy()
        p, q0 = tot[:m], tot[m:2 * m]
        SY = tot[2 * m:2 * m + m * m].reshape(m, m)
        YY = tot[2 * m + m * m:2 * m + 2 * m * m].reshape(m, m)
        # --- K2 (every rank solves redundantly), slots in ring order 0..m-1
        R = np.triu(SY)
        gamma = SY[m - 1, m - 1] / YY[m - 1, m - 1]
        u = np.linalg.solve(R, p)
        a = np.linalg.solve(R.T, np.diag(R) * u + gamma * (YY @ u) - gamma * q0)
        # --- K3 on the shard
        d_loc = gamma * g_loc + a @ S[:, sl] + (-gamma * u) @ Y[:, sl]

        # unsharded oracle two-loop on the full vectors
        from oracle import stochqn_np as O
        mem = O.BfgsMem(m, n, 0.0, 0.0, 1, np.float64)
        mem.s_mem[:] = S
        mem.y_mem[:] = Y
        mem.mem_used, mem.mem_st_ix = m, 0
        d_ref = g_full.copy()
        O.approx_inv_hess_grad(d_ref, None, 0.0, mem, 0)
        err = np.max(np.abs(d_loc - d_ref[sl])) / np.max(np.abs(d_ref))
        q.put((rank, float(err)))
    finally:
        dist.destroy_process_group()


def test_sharded_step_equals_unsharded_oracle_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 1001, 4, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=10) for _ in range(world))
    assert [r for r, _ in got] == [0, 1]
    assert all(err < 1e-10 for _, err in got), got
