"""Worker of tests/test_gpu_zz_inprocess.py: the peer-memory exchange kernels on ONE GPU.

`stochqn_b200_comm_init_inprocess` makes world_size communicators whose ranks all live in this process; every rank gets
its own non-blocking stream and the calls of one collective are issued rank after rank without a host synchronisation
in between (the waiting kernels of all ranks run together).  The kernels are the ones the one-process-per-GPU case
runs; only the way the peers' buffers are mapped differs (plain pointers instead of cudaIpc).

Launched with CUDA_MODULE_LOADING=EAGER (with lazy loading the first launch of a kernel may have to wait for the device
to drain, and a rank-0 kernel that is polling for rank 1 never drains while rank 1's launch sits behind that load) and
CUDA_DEVICE_MAX_CONNECTIONS=32 (one hardware queue per rank stream: on a shared queue rank 1's launch would sit behind
the kernel of rank 0 that depends on rank 0's polling kernel).  Both are hazards of sharing ONE device.

argv: out_json
"""
import ctypes as C
import json
import os
import sys
import traceback

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

from oracle import multinomial_np as MN                 # noqa: E402
from stochqn_b200 import _lib                           # noqa: E402


class _Raw:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def view(ptr, n, tdt):
    return torch.as_tensor(_Raw(ptr, n, "<f8" if tdt == torch.float64 else "<f4"), device="cuda")


class Group:
    def __init__(self, abi, world):
        self.abi, self.lib, self.world = abi, abi.lib, world
        arr = (C.c_void_p * world)()
        rc = self.lib.stochqn_b200_comm_init_inprocess(world, arr)
        assert rc == 0, (rc, _lib.last_error(abi))
        self.comms = [C.c_void_p(arr[r]) for r in range(world)]
        self.streams = [torch.cuda.Stream() for _ in range(world)]

    def st(self, r):
        return C.c_void_p(self.streams[r].cuda_stream)

    def check(self):
        torch.cuda.synchronize()
        for r in range(self.world):
            assert self.lib.stochqn_b200_comm_error(self.comms[r]) == 0, "rank %d: an exchange timed out" % r

    def close(self):
        torch.cuda.synchronize()
        for c in self.comms:
            self.lib.stochqn_b200_comm_destroy(c)


def seq_sum(parts):
    """sum in rank order, in the parts' own dtype (what the kernels do)"""
    acc = parts[0].clone()
    for p in parts[1:]:
        acc = acc + p
    return acc


def case_allreduce(abi, world, count, rounds=4):
    g = Group(abi, world)
    gen = torch.Generator(device="cuda").manual_seed(world * 1000 + count)
    for k in range(rounds):
        bufs = [torch.randn(count, device="cuda", dtype=torch.float64, generator=gen) * (10.0 ** (r % 5)) for r in range(world)]
        want = torch.zeros(count, device="cuda", dtype=torch.float64)
        for b in bufs:                                     # the kernel starts from 0 and adds rank 0, 1, ...
            want = want + b
        torch.cuda.synchronize()
        for r in range(world):
            rc = g.lib.stochqn_b200_allreduce_f64(g.comms[r], bufs[r].data_ptr(), count, g.st(r))
            assert rc == 0, (rc, _lib.last_error(abi))
        g.check()
        for r in range(world):
            assert torch.equal(bufs[r], want), "round %d rank %d: max diff %g" % (k, r, float((bufs[r] - want).abs().max()))
    g.close()
    return {"rounds": rounds}


def case_allgather(abi, tdt, world, blk, rounds=3):
    g = Group(abi, world)
    gen = torch.Generator(device="cuda").manual_seed(world * 7 + blk)
    for k in range(rounds):
        blocks = [torch.randn(blk, device="cuda", dtype=tdt, generator=gen) for _ in range(world)]
        want = torch.cat(blocks)
        torch.cuda.synchronize()
        ptrs = []
        for r in range(world):
            gp = C.c_void_p()
            rc = g.lib.stochqn_b200_all_gather_p2p(g.comms[r], blocks[r].data_ptr(), blk, C.byref(gp), g.st(r))
            assert rc == 0, (rc, _lib.last_error(abi))
            ptrs.append(gp.value)
        g.check()
        assert len(set(ptrs)) == world, "every rank owns its gathered vector"
        for r in range(world):
            got = view(ptrs[r], world * blk, tdt)
            assert torch.equal(got, want), "round %d rank %d" % (k, r)
    g.close()
    return {"rounds": rounds}


def case_reduce_scatter(abi, tdt, world, blk, rounds=4):
    """pull reduce-scatter: rounds alternate between the library's send vector (the producer writes into it) and a
    vector that lives elsewhere (copied in)"""
    g = Group(abi, world)
    gen = torch.Generator(device="cuda").manual_seed(world * 13 + blk)
    n = world * blk
    for k in range(rounds):
        full = [torch.randn(n, device="cuda", dtype=tdt, generator=gen) * (1.0 + r) for r in range(world)]
        outs = [torch.full((blk,), float("nan"), device="cuda", dtype=tdt) for _ in range(world)]
        want = seq_sum(full)
        send = []
        for r in range(world):
            if k % 2 == 0:
                sp = C.c_void_p()
                rc = g.lib.stochqn_b200_p2p_send_buffer(g.comms[r], blk, C.byref(sp))
                assert rc == 0, (rc, _lib.last_error(abi))
                view(sp.value, n, tdt).copy_(full[r])
                send.append(sp.value)
            else:
                send.append(full[r].data_ptr())
        torch.cuda.synchronize()
        for r in range(world):
            rc = g.lib.stochqn_b200_reduce_scatter_p2p(g.comms[r], send[r], outs[r].data_ptr(), blk, g.st(r))
            assert rc == 0, (rc, _lib.last_error(abi))
        g.check()
        if k % 2 == 0:                                     # double-buffered: the next call reads the other vector
            for r in range(world):
                sp = C.c_void_p()
                assert g.lib.stochqn_b200_p2p_send_buffer(g.comms[r], blk, C.byref(sp)) == 0
                assert sp.value != send[r]
        for r in range(world):
            w = want[r * blk:(r + 1) * blk]
            assert torch.equal(outs[r], w), "round %d rank %d: max diff %g" % (k, r, float((outs[r] - w).abs().max()))
    g.close()
    return {"rounds": rounds}


def _mn_problem(tdt, world, rows_per_rank, d, K, seed):
    gen = torch.Generator(device="cuda").manual_seed(seed)
    B = world * rows_per_rank
    X = torch.randn(B, d, device="cuda", dtype=tdt, generator=gen) / d ** 0.5
    lab = torch.randint(0, K, (B,), device="cuda", generator=gen).to(torch.int32)
    w = torch.randn(K * (d + 1), device="cuda", dtype=tdt, generator=gen) * 0.3
    sw = torch.full((B,), 1.0 / B, device="cuda", dtype=tdt)
    return X, lab, w, sw


def _oracle_grad(X, lab, w, sw, d, K, alpha):
    Y = np.zeros((X.shape[0], K))
    Y[np.arange(X.shape[0]), lab.cpu().numpy()] = 1.0
    _, g, _ = MN.multinomial_loss_grad(w.double().cpu().numpy(), X.double().cpu().numpy(), Y, alpha, sw.double().cpu().numpy())
    return g


def case_rowsharded_gradient(abi, tdt, world, how, rows_per_rank=256, d=1024, K=512, rounds=3):
    """One mini-batch gradient with the batch rows sharded over the ranks of an in-process group, combined by
    how = "pull"  : all_gather_p2p of the point, multinomial_loss_grad into the library's send vector, reduce_scatter_p2p
    how = "fused" : all_gather_p2p of the point, multinomial_grad_reduce_scatter (reduce-scatter inside the GEMM epilogue)
    against the oracle's gradient on the union of the rows."""
    g = Group(abi, world)
    lib = g.lib
    n = K * (d + 1)
    assert n % world == 0
    blk = n // world
    alpha = 1e-3
    esz = 8 if tdt == torch.float64 else 4
    work = [torch.empty(lib.stochqn_b200_multinomial_work_size(rows_per_rank, d, K), device="cuda", dtype=torch.uint8) for _ in range(world)]
    errs = []
    for k in range(rounds):
        X, lab, w, sw = _mn_problem(tdt, world, rows_per_rank, d, K, 50 + k)
        want = _oracle_grad(X, lab, w, sw, d, K, alpha)
        outs = [torch.full((blk,), float("nan"), device="cuda", dtype=tdt) for _ in range(world)]
        torch.cuda.synchronize()
        # two passes over the ranks: the first call of a collective allocates its buffers for the whole group and drains the
        # device to do so - it must not find rank 0's all-gather waiting for a rank 1 that has not been issued yet
        points = []
        for r in range(world):
            gp = C.c_void_p()
            rc = lib.stochqn_b200_all_gather_p2p(g.comms[r], w.data_ptr() + r * blk * esz, blk, C.byref(gp), g.st(r))
            assert rc == 0, (rc, _lib.last_error(abi))
            points.append(gp.value)
        for r in range(world):
            r0 = r * rows_per_rank
            args = (X.data_ptr() + r0 * d * esz, d, None, K, lab.data_ptr() + r0 * 4, sw.data_ptr() + r0 * esz, rows_per_rank, d, K, 1, points[r], alpha / world)
            if how == "fused":
                rc = lib.stochqn_b200_multinomial_grad_reduce_scatter(g.comms[r], *args, outs[r].data_ptr(), blk, work[r].data_ptr(), g.st(r))
                assert rc == 0, (rc, _lib.last_error(abi))
            else:
                sp = C.c_void_p()
                rc = lib.stochqn_b200_p2p_send_buffer(g.comms[r], blk, C.byref(sp))
                assert rc == 0, (rc, _lib.last_error(abi))
                rc = lib.stochqn_b200_multinomial_loss_grad(*args, sp.value, None, work[r].data_ptr(), g.st(r))
                assert rc == 0, (rc, _lib.last_error(abi))
                rc = lib.stochqn_b200_reduce_scatter_p2p(g.comms[r], sp.value, outs[r].data_ptr(), blk, g.st(r))
                assert rc == 0, (rc, _lib.last_error(abi))
        g.check()
        got = torch.cat(outs).double().cpu().numpy()
        errs.append(float(np.max(np.abs(got - want)) / np.max(np.abs(want))))
    g.close()
    tol = 1e-11 if tdt == torch.float64 else 3e-3          # fp32: tf32 tensor cores (tests/test_gpu_multinomial.py states the bound)
    assert max(errs) <= tol, errs
    return {"rel_err": max(errs), "tol": tol}


def main():
    out = sys.argv[1]
    res = {}

    timed_out = set()

    def run(name, fn, *a, **kw):
        kind = name.split("_")[0]
        if kind in timed_out:                              # every further case of the kind would wait its 20 s as well
            res[name] = {"ok": False, "error": "not run: an earlier %s case timed out" % kind}
            json.dump(res, open(out, "w"), indent=1)
            return
        try:
            r = fn(*a, **kw) or {}
            r["ok"] = True
        except Exception as e:                             # noqa: BLE001
            r = {"ok": False, "error": "%s: %s" % (type(e).__name__, e), "trace": traceback.format_exc()[-1500:]}
            if "timed out" in str(e):
                timed_out.add(kind)
        res[name] = r
        json.dump(res, open(out, "w"), indent=1)

    abis = {"f64": (_lib.load(np.float64), torch.float64), "f32": (_lib.load(np.float32), torch.float32)}
    for world in (2, 3, 8):
        for count in (1, 130, 2048):
            run("allreduce_w%d_c%d" % (world, count), case_allreduce, abis["f64"][0], world, count)
    for tag, (abi, tdt) in abis.items():
        for world in (2, 4, 8):
            for blk in (1024, 1001):
                run("allgather_%s_w%d_b%d" % (tag, world, blk), case_allgather, abi, tdt, world, blk)
        for world in (2, 3, 4, 8):
            for blk in (4096, 1001, 300000):
                run("reduce_scatter_%s_w%d_b%d" % (tag, world, blk), case_reduce_scatter, abi, tdt, world, blk)
    # more than 128 rows per rank: the one-launch small-batch gradient is a cooperative grid sized for an empty GPU and
    # would wait for the other rank's polling kernel to leave - a hazard of sharing ONE device, not of the exchange
    run("rowsharded_pull_f64_w2", case_rowsharded_gradient, *abis["f64"], 2, "pull", rows_per_rank=160, d=60, K=40)
    run("rowsharded_pull_f64_w4", case_rowsharded_gradient, *abis["f64"], 4, "pull", rows_per_rank=144, d=63, K=40)
    run("rowsharded_pull_f32_w2", case_rowsharded_gradient, *abis["f32"], 2, "pull")
    run("rowsharded_pull_f32_w4", case_rowsharded_gradient, *abis["f32"], 4, "pull")
    run("rowsharded_fused_f32_w2", case_rowsharded_gradient, *abis["f32"], 2, "fused")
    run("rowsharded_fused_f32_w4", case_rowsharded_gradient, *abis["f32"], 4, "fused")


if __name__ == "__main__":
    main()
