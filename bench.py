#!/usr/bin/env python
"""bench.py - optimizer steps/s of free-mode oLBFGS on the chained Rosenbrock function.

Workload = BASELINE.json config 4: fp64, n = 2^27 parameters, mem_size 10,
hess_init 0 (gamma scaling), y_reg 0, min_curvature 1e-4, check_nan 1, step 1e-4,
x0[i] = 0.95 + 1e-4*((uint32)(i*2654435761) mod 1000).  One "step" = one oLBFGS iteration =
serve calc_grad -> run_oLBFGS (take the step) -> serve calc_grad_same_batch -> run_oLBFGS (pair).

    python bench.py [--gpus N] [--steps K] [--warmup W]            this repo (CUDA, sm_100a)
    python bench.py --impl reference [...]                          the reference's own CPU build (oracle/_ref)

N > 1: launched by torchrun, one rank per GPU; the parameter vector (total length fixed at n: strong
scaling) shards by contiguous blocks, dot partials go through one small exchange per phase.

Timed region (both arms): the pair memory is FULL when the clock starts - max(W, mem_size + 2) untimed
iterations are run first and mem_used == mem_size is asserted - so every timed iteration moves the same
(4*mem_size + 14) n-vectors.  This arm then times BLOCKS blocks of K iterations each (barrier + synchronize on
both sides of every block, CUDA events, max over ranks per block) and reports the MEDIAN block; the spread is
printed beside it.

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (x, gradients, pairs all in HBM,
gradient requests served by the bundled device callback).  `e2e` = the same metric through the drop-in
C ABI with HOST buffers: x and grad live in pinned host memory, the gradient is evaluated by a host
(C + OpenMP) callback as a user of the reference would, and every call stages host<->device copies.
`check` = what was verified in this very run: the probes of the iterate after warm-up + K iterations (the same
iteration count the reference arm runs), their distance from the reference C library's on the same workload
(`vs_reference*`), and for N > 1 the sharded paths against the unsharded one on rank 0 (`sharded_parity`).
`secondary` = BASELINE configs 1, 2, 3, 5 (N = 1) / config 5 row-sharded (N > 1), each with its own CPU reference.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_DEFAULT = 2 ** 27
MEM = 10
STEP = 1e-4
MIN_CURV = 1e-4
BLOCKS = 5
METRIC = "optimizer steps/s at n=2^27, m=10 (oLBFGS, fp64, chained Rosenbrock)"
REF_MARCH = "x86-64-v3"          # oracle/build_ref.py: the artefact travels to the GPU box, so not -march=native


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def effective_warmup(w):
    """Untimed iterations before the clock starts: the pair memory must be full (mem_size accepted pairs) and the
    first full-memory iteration (new K1 instantiation, pending Gram column) must be behind us."""
    return max(int(w), MEM + 2, 3)


def config_dict(n, gpus):
    """The SAME dictionary from both arms (the driver compares them)."""
    return {"workload": "BASELINE config 4: free-mode oLBFGS, chained Rosenbrock, n=2^27, mem_size=10, fp64",
            "n": int(n), "n_gpus": int(gpus), "mem_size": MEM, "step": STEP, "min_curvature": MIN_CURV, "check_nan": 1,
            "hess_init": 0.0, "y_reg": 0.0, "memory_full_at_start": True,
            "l2": "inputs larger than L2 (%.1f GiB streamed per step)" % ((4 * MEM + 14) * n * 8 / 2 ** 30)}


def probe_indices(n):
    return [0, 1 if n > 1 else 0, n // 4, n // 2, (3 * n) // 4, n - 2 if n > 1 else 0, n - 1]


def rel_probe_distance(a, b):
    """Largest relative difference over x_norm, x_sum and the probe entries of two probe records."""
    worst = 0.0
    for k in ("x_norm", "x_sum"):
        if k in a and k in b and b[k] != 0:
            worst = max(worst, abs(a[k] - b[k]) / abs(b[k]))
    for u, v in zip(a.get("probes", []), b.get("probes", [])):
        worst = max(worst, abs(u - v) / max(abs(v), 1e-300))
    return worst


def _cache_paths():
    # (a file next to the sources would travel with the snapshot and go stale: the temporary directory of the box only)
    return [os.path.join(tempfile.gettempdir(), "stochqn_b200_reference_arm.json")]


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own C library driven by oracle/rosen_harness.c
# ------------------------------------------------------------------------------------------------
def _harness_path():
    return os.path.join(ROOT, "oracle", "_ref", "rosen_harness_f64")


def _run_harness(n, warmup, steps, threads):
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), OPENBLAS_NUM_THREADS=str(threads))
    out = subprocess.run([_harness_path(), str(n), str(MEM), str(warmup), str(steps), str(threads), str(MIN_CURV), str(STEP), "1"],
                         capture_output=True, text=True, env=env, check=True)
    return json.loads(out.stdout.strip().splitlines()[-1])


def cpu_reference_run(steps, warmup, budget_s, n_target=N_DEFAULT):
    """Time the reference on the host cores on a bounded sample of the workload: the same problem at the
    largest power-of-two length n_s <= n_target whose (warmup + steps) iterations fit `budget_s`, with the
    better of {1, all} threads; throughput is scaled by n_s / n_target (every operation of the step is a
    streaming pass, cost linear in n).  `warmup` >= mem_size + 2, so the timed iterations run with the memory full."""
    if not os.path.exists(_harness_path()):
        return None
    cores = os.cpu_count() or 1
    cal_n = 2 ** 22
    best = None
    for th in sorted({1, cores}):
        r = _run_harness(cal_n, MEM + 2, 3, th)
        if best is None or r["steps_per_s"] > best[1]["steps_per_s"]:
            best = (th, r)
    th, r = best
    per_elem_step = 1.0 / (r["steps_per_s"] * cal_n)                 # seconds per element per iteration
    n_s = n_target
    while n_s > 2 ** 16 and per_elem_step * n_s * (warmup + steps) * 1.3 > budget_s:
        n_s //= 2
    rr = _run_harness(n_s, warmup, steps, th)
    scale = n_s / float(n_target)
    probes = {k: rr[k] for k in ("x_norm", "x_sum", "probes", "probe_idx", "info_events", "mem_used", "mem_st_ix", "niter") if k in rr}
    probes.update(n=n_s, iterations=warmup + steps)
    return dict(value=rr["steps_per_s"] * scale, opt_only=rr["opt_steps_per_s"] * scale, cores=th, host_cores=cores, n_sample=n_s,
                raw_steps_per_s=rr["steps_per_s"], seconds=rr["seconds"], steps=steps, warmup=warmup, probes=probes)


def _sample_text(r):
    return ("reference C library (unmodified src/stochqn.c, gcc -O2 -fopenmp -march=%s [the reference's own flags say -march=native; "
            "only its OpenMP element-wise loops are affected, BLAS dispatches at run time], SciPy's OpenBLAS) on the same oLBFGS/Rosenbrock "
            "workload at n=%d (1/%d of 2^27), %d warm-up (memory full) + %d timed iterations, %d thread(s) of %d host cores; "
            "steps/s scaled by n_sample/2^27 (all passes are linear in n)"
            % (REF_MARCH, r["n_sample"], N_DEFAULT // r["n_sample"], r["warmup"], r["steps"], r["cores"], r["host_cores"]))


def main_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    t0 = time.time()
    warm = effective_warmup(args.warmup)
    r = cpu_reference_run(args.steps, warm, budget_s=args.ref_budget, n_target=args.n)
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/rosen_harness_f64 missing (reference not built)"}))
        return 0
    sample = _sample_text(r)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "warmup_run": warm, "ms_per_step": 1e3 / r["value"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args.n, args.gpus),
        "cpu_baseline": {"value": r["value"], "unit": "steps/s", "cores": r["cores"], "kind": "reference", "sample": sample},
        "e2e": {"value": r["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "check": r["probes"], "wall_s": time.time() - t0,
    }
    for p in _cache_paths():            # the CUDA arm, run next on the same box, compares its own probes with these
        try:
            json.dump(r["probes"], open(p, "w"))
        except OSError:
            pass
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "50"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, reasons, mx = [], set(), None
        for ln in self.f.read().splitlines():
            parts = [s.strip() for s in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx = float(parts[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------
# this repo's arm: one device-resident oLBFGS / Rosenbrock run (sharded when a communicator is given)
# ------------------------------------------------------------------------------------------------
class CudaRosen:
    """Free-mode oLBFGS on the chained Rosenbrock function through the C ABI with device pointers; the parameter
    vector is the block [offset, offset + n_local) of a vector of length n when `comm` is given."""

    def __init__(self, torch, abi, n, rank=0, world=1, comm=None, mem=MEM, trace=False, min_curv=MIN_CURV):
        from stochqn_b200 import _lib
        from stochqn_b200.distributed import shard_bounds

        self.torch, self.abi, self.lib, self._lib = torch, abi, abi.lib, _lib
        self.n, self.rank, self.world, self.comm, self.mem = n, rank, (world if comm is not None else 1), comm, mem
        self.offset, self.n_local = shard_bounds(n, rank, self.world) if comm is not None else (0, n)
        lib = self.lib
        self.stream = torch.cuda.current_stream().cuda_stream
        self.x = torch.empty(self.n_local, device="cuda", dtype=torch.float64)
        self.g = torch.empty(self.n_local, device="cuda", dtype=torch.float64)
        self.g2 = torch.empty(self.n_local, device="cuda", dtype=torch.float64)
        self.prefetched, self.halo_valid = False, False
        self.halo = torch.zeros(2, device="cuda", dtype=torch.float64)
        self.scratch = torch.zeros(2 * max(self.world, 1), device="cuda", dtype=torch.float64)
        lib.stochqn_b200_rosenbrock_x0(self.x.data_ptr(), self.n_local, self.offset, self.stream)
        self.ws = lib.initialize_oLBFGS(self.n_local, mem, 0.0, 0.0, min_curv, 1, 1)
        if not self.ws:
            raise SystemExit("initialize_oLBFGS failed: " + _lib.last_error(abi))
        if comm is not None:
            assert lib.stochqn_b200_set_comm(self.ws, comm, n) == 0
        self.req, self.task, self.info = C.c_void_p(), C.c_int(), C.c_int()
        self.info_events = 0
        self.trace = [] if trace else None
        self._call(self.g)                                                # section 0: first request

    def _call(self, gbuf):
        ret = self.lib.run_oLBFGS(STEP, self.x.data_ptr(), gbuf.data_ptr(), C.byref(self.req), C.byref(self.task), self.ws, C.byref(self.info))
        self.info_events += self.info.value != 200
        if self.trace is not None:
            w = self.ws.contents
            m = w.bfgs_memory.contents
            self.trace.append((int(self.task.value), int(ret), int(self.info.value), int(w.niter), int(w.section), int(m.mem_used), int(m.mem_st_ix)))

    def _serve(self, gbuf, same_point=False):
        """The caller's gradient callback (bundled device kernel) at *req.  Sharded: the halo exchange is fused into the
        gradient kernel (peer memory), one launch; a second evaluation at the SAME point reuses the neighbours' edge values
        the first one fetched (no second exchange)."""
        lib = self.lib
        if self.comm is not None and self.world > 1 and not (same_point and self.halo_valid):
            lib.stochqn_b200_rosenbrock_grad_sharded(self.req.value, gbuf.data_ptr(), self.n_local, self.offset, self.n, self.rank, self.world,
                                                     self.comm, self.halo.data_ptr(), self.scratch.data_ptr(), self.stream)
            self.halo_valid = True
        else:
            lib.stochqn_b200_rosenbrock_grad(self.req.value, gbuf.data_ptr(), self.n_local, self.offset, self.n, self.halo.data_ptr(), self.stream)

    def iteration(self):
        """One oLBFGS iteration through the free-mode ABI.  The two gradients of an iteration live in two buffers, so that
        the gradient the NEXT request will ask for (calc_grad at x, which the pair call does not move: stochqn.c:1024-1031)
        is already queued behind the same-batch gradient when the pair call waits for its curvature flag - the GPU never
        idles between the pair call's return and the caller's next launch.  Same requests, same evaluations, same order
        of optimizer calls; a rejected step simply discards nothing (its next request is served when it is made)."""
        if not self.prefetched:
            self._serve(self.g)                 # calc_grad
        self._call(self.g)                      # step
        self.prefetched = False
        if self.task.value == 102:
            self._serve(self.g2)                # calc_grad_same_batch at the new x
            self._serve(self.g, same_point=True)   # what the pair call will request next: calc_grad at the same x
            self.prefetched = True
            self._call(self.g2)                 # pair
            assert self.task.value == 101 and self.req.value == self.x.data_ptr()

    def run(self, iters):
        for _ in range(iters):
            self.iteration()

    def mem_used(self):
        return int(self.ws.contents.bfgs_memory.contents.mem_used)

    def probes(self, dist=None):
        """x_norm, x_sum and the probe entries of the GLOBAL iterate (collected over the ranks when sharded)."""
        torch = self.torch
        idx = probe_indices(self.n)
        v = torch.zeros(2 + len(idx), device="cuda", dtype=torch.float64)
        v[0] = torch.sum(self.x * self.x)
        v[1] = torch.sum(self.x)
        for k, i in enumerate(idx):
            if self.offset <= i < self.offset + self.n_local:
                v[2 + k] = self.x[i - self.offset]
        if dist is not None and self.comm is not None and self.world > 1:
            dist.all_reduce(v)
        h = v.cpu().tolist()
        w = self.ws.contents
        m = w.bfgs_memory.contents
        return {"n": self.n, "iterations": int(w.niter), "x_norm": h[0] ** 0.5, "x_sum": h[1], "probes": h[2:], "probe_idx": idx,
                "info_events": int(self.info_events), "mem_used": int(m.mem_used), "mem_st_ix": int(m.mem_st_ix), "niter": int(w.niter)}

    def gather_x(self, dist):
        torch = self.torch
        if self.comm is None or self.world <= 1:
            return self.x.clone()
        from stochqn_b200.distributed import shard_bounds
        parts = [torch.empty(shard_bounds(self.n, r, self.world)[1], device="cuda", dtype=torch.float64) for r in range(self.world)]
        dist.all_gather(parts, self.x)
        return torch.cat(parts)

    def close(self):
        if self.ws:
            self.lib.dealloc_oLBFGS(self.ws)
            self.ws = None
        self.x = self.g = self.g2 = None


def sharded_parity(torch, dist, abi, rank, world, comm_default):
    """N > 1, before anything is timed: the SAME small problem (n = 100 003: not a multiple of the rank count) run
    (a) sharded over all ranks with the default exchange path (peer-memory mailboxes inside the kernels),
    (b) sharded with the ncclAllReduce path (a second communicator created with STOCHQN_B200_NO_P2P=1),
    (c) unsharded on rank 0.  Task / return / info / counter sequences must be identical on every rank and equal to
    the unsharded ones; the gathered iterate must agree with the unsharded one to 1e-10 (relative, max norm)."""
    from stochqn_b200.distributed import init_comm

    lib = abi.lib
    n_s, iters = 100003, 30
    res = {"n": n_s, "iterations": iters, "world": world}
    ref_x, ref_trace = None, None
    if rank == 0:
        r0 = CudaRosen(torch, abi, n_s, mem=5, trace=True)
        r0.run(iters)
        torch.cuda.synchronize()
        ref_x, ref_trace = r0.x.clone(), list(r0.trace)
        r0.close()
    dist.barrier()
    old = os.environ.get("STOCHQN_B200_NO_P2P")
    os.environ["STOCHQN_B200_NO_P2P"] = "1"
    comm_nccl = init_comm(abi, rank, world)
    if old is None:
        del os.environ["STOCHQN_B200_NO_P2P"]
    else:
        os.environ["STOCHQN_B200_NO_P2P"] = old
    ok_all = True
    for name, cm in (("default", comm_default), ("nccl", comm_nccl)):
        run = CudaRosen(torch, abi, n_s, rank, world, cm, mem=5, trace=True)
        run.run(iters)
        torch.cuda.synchronize()
        xs = run.gather_x(dist)
        traces = [None] * world
        dist.all_gather_object(traces, run.trace)
        entry = {"uses_p2p": int(lib.stochqn_b200_comm_uses_p2p(cm)), "calls": len(run.trace)}
        if rank == 0:
            same = all(t == traces[0] for t in traces)
            match = traces[0] == ref_trace
            err = float((xs - ref_x).abs().max().item() / ref_x.abs().max().item())
            entry.update(sequences_identical_on_all_ranks=same, sequences_equal_unsharded=match, x_rel_err_vs_unsharded=err,
                         pairs=int(run.mem_used()), ok=bool(same and match and err <= 1e-10))
            ok_all = ok_all and entry["ok"]
        res[name] = entry
        run.close()
        dist.barrier()
    lib.stochqn_b200_comm_destroy(comm_nccl)
    res["ok"] = bool(ok_all)
    return res


def main_b200(args):
    import numpy as np
    import torch

    from stochqn_b200 import _lib
    from stochqn_b200.distributed import init_comm

    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    abi = _lib.load(np.float64)
    lib = abi.lib
    n = args.n
    comm = init_comm(abi, rank, world) if world > 1 else None
    check = {}

    def fence():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- multi-GPU correctness, before timing -------------------------------------------------------------
    if world > 1 and not args.no_checks:
        check["sharded_parity"] = sharded_parity(torch, dist, abi, rank, world, comm)
        fence()

    # ---- the timed run ------------------------------------------------------------------------------------
    warm = effective_warmup(args.warmup)
    run = CudaRosen(torch, abi, n, rank, world, comm)
    n_local = run.n_local
    sampler = ClockSampler(local_rank) if rank == 0 else None      # samples the warm-up and the timed blocks (both under load)
    run.run(warm)
    fence()
    assert run.mem_used() == MEM, "pair memory not full after %d warm-up iterations (mem_used = %d)" % (warm, run.mem_used())
    lib.stochqn_b200_set_option(run.ws, _lib.OPT_PROFILE, 1)
    block_ms, probes_first, launches = [], None, 0
    for b in range(BLOCKS):
        fence()
        launches0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run.run(args.steps)
        e1.record()
        fence()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        block_ms.append(float(t.item()))
        if b == 0:
            launches = _lib.launch_count() - launches0
            probes_first = run.probes(dist)       # iterate after warm + K iterations: what the reference arm ends with
        assert run.mem_used() == MEM
    clocks = sampler.stop() if sampler else None
    srt = sorted(block_ms)
    ms = srt[len(srt) // 2]
    value = args.steps / (ms * 1e-3)
    st = {k: _lib.get_stat(abi, run.ws, v) for k, v in dict(k1=1, k1n=2, k3=3, k3n=4, k4=5, k4n=6).items()}
    lib.stochqn_b200_set_option(run.ws, _lib.OPT_PROFILE, 0)
    probes_end = run.probes(dist)
    run.close()

    # "min_curvature = 1e-4 (R / Python default; also report 0)" - the same loop with the curvature test off
    # (the reference then allocates no backup buffers; here the test is a host decision on two dots K4 computes anyway)
    mc0 = None
    if not args.no_checks:
        r0 = CudaRosen(torch, abi, n, rank, world, comm, min_curv=0.0)
        r0.run(warm)
        fence()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r0.run(args.steps)
        e1.record()
        fence()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t = float(t.item())
        mc0 = {"value": args.steps / (t * 1e-3), "unit": "steps/s", "ms_per_step": t / args.steps, "blocks": 1, "mem_used": r0.mem_used(),
               "info_events": int(r0.info_events)}
        r0.close()
    check.update(after_first_block=probes_first, after_all_blocks={k: probes_end[k] for k in ("iterations", "x_norm", "info_events", "mem_used")},
                 x_norm=probes_end["x_norm"], info_events=probes_end["info_events"], mem_used=probes_end["mem_used"])

    # ---- roofline of the dominant kernel (K3: fused combine + update), per launch ------------------
    # every timed launch ran with mem_used == MEM (asserted above), so bytes per launch are constant
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    vec_bytes = n_local * 8
    k3_ms = st["k3"] / max(st["k3n"], 1)
    k1_ms = st["k1"] / max(st["k1n"], 1)
    k4_ms = st["k4"] / max(st["k4n"], 1)
    k3_bytes = (2 * MEM + 4) * vec_bytes               # read g, S, Y, x; write x, s_new  (SURVEY 8(d)); the grad write-back is not counted
    k1_bytes = (2 * MEM + 2) * vec_bytes               # read g, S, Y; write grad_prev
    k4_bytes = 4 * vec_bytes
    traffic, traffic_note = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = float(tj["k3_dram_bytes_per_launch"]) * n_local / float(tj["n_local"])
        traffic_note = "dram__bytes_read+write of one K3 launch captured at n_local=%d, scaled by n_local/%d (%s)" % (
            int(tj["n_local"]), int(tj["n_local"]), tj.get("source", "profiles/"))
    except Exception:
        pass
    step_bytes = (4 * MEM + 14) * vec_bytes * world
    roofline = {"bound": "hbm", "kernel": "k3_combine (fused combine + x update + new s)", "achieved": k3_bytes / k3_ms / 1e6,
                "peak": peak, "unit": "GB/s", "frac": k3_bytes / k3_ms / 1e6 / peak, "traffic": traffic, "traffic_note": traffic_note,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth, burst)" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                "per_launch_bytes": k3_bytes, "avg_launch_ms": k3_ms, "launches_timed": int(st["k3n"]), "pairs_at_every_launch": MEM,
                "other_kernels": {
                    "k1_dots": {"achieved": k1_bytes / k1_ms / 1e6, "frac": k1_bytes / k1_ms / 1e6 / peak, "avg_launch_ms": k1_ms, "per_launch_bytes": k1_bytes},
                    "k4_pair": {"achieved": k4_bytes / k4_ms / 1e6, "frac": k4_bytes / k4_ms / 1e6 / peak, "avg_launch_ms": k4_ms, "per_launch_bytes": k4_bytes}},
                "step": {"algorithmic_bytes": step_bytes, "achieved": step_bytes / (ms / args.steps) / 1e6,
                         "frac_of_aggregate_peak": step_bytes / (ms / args.steps) / 1e6 / (peak * world),
                         "frac_of_aggregate_8TBs": step_bytes / (ms / args.steps) / 1e6 / (8000.0 * world)}}

    # ---- end-to-end through the drop-in C ABI with HOST buffers --------------------------------------
    e2e = None if args.no_e2e else run_e2e(args, lib, abi, rank, world, local_rank, n, comm, dist, torch, np)

    # ---- CPU baseline (rank 0, N = 1 only) + direct comparison of the two arms on its sample ---------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(8, MEM + 2, budget_s=25.0, n_target=n)
        if r is not None:
            cpu = {"value": r["value"], "unit": "steps/s", "cores": r["cores"], "kind": "reference", "sample": _sample_text(r),
                   "optimizer_only_value": r["opt_only"]}
            # the same n and iteration count on the GPU: probes of the two final iterates side by side
            rr = CudaRosen(torch, abi, r["n_sample"])
            rr.run(r["warmup"] + r["steps"])
            mine = rr.probes()
            rr.close()
            check["vs_reference_sample"] = {"n": r["n_sample"], "iterations": r["warmup"] + r["steps"], "rel": rel_probe_distance(mine, r["probes"]),
                                            "counters_equal": all(mine[k] == r["probes"].get(k) for k in ("info_events", "mem_used", "mem_st_ix", "niter")),
                                            "x_norm": [mine["x_norm"], r["probes"]["x_norm"]]}
    # ---- the reference arm's probes (it ran first on this box): same n, same iteration count? ---------------
    if rank == 0:
        for p in _cache_paths():
            try:
                ref = json.load(open(p))
            except Exception:
                continue
            if ref.get("n") == n and ref.get("iterations") == probes_first["iterations"]:
                check["vs_reference_rel"] = rel_probe_distance(probes_first, ref)
                check["vs_reference_counters_equal"] = all(probes_first[k] == ref.get(k) for k in ("info_events", "mem_used", "mem_st_ix", "niter"))
            else:
                check["vs_reference_note"] = "reference arm ran n=%s, %s iterations; this arm n=%d, %d after the first block: not comparable" % (
                    ref.get("n"), ref.get("iterations"), n, probes_first["iterations"])
            break

    # ---- the other BASELINE configurations ----------------------------------------------------------------
    secondary = None
    if not args.no_secondary:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import secondary_configs
            secondary = secondary_configs.run(rank, world, comm, dist, peaks, budget_s=args.secondary_budget)
        except Exception as e:          # the headline line must not be lost to a secondary failure
            secondary = {"error": "%s: %s" % (type(e).__name__, e)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "warmup_run": warm, "blocks": BLOCKS, "block_ms": block_ms, "block_spread": (srt[-1] - srt[0]) / ms,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config_dict(n, world),
            "run": {"n_per_gpu": n_local, "grad_writeback": 1, "value_is": "median of %d blocks of %d iterations, each max over ranks" % (BLOCKS, args.steps),
                    "callbacks": "bundled device Rosenbrock gradient (halo exchange fused in when sharded)"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "check": check, "min_curvature_0": mc0, "secondary": secondary,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        if comm is not None:
            lib.stochqn_b200_comm_destroy(comm)
        dist.destroy_process_group()
    return 0


def run_e2e(args, lib, abi, rank, world, local_rank, n, comm, dist, torch, np):
    """Same metric through run_oLBFGS with HOST pointers: x / grad in pinned host memory, gradient by a host
    (C + OpenMP) callback, H2D / D2H staging inside every call (the library's compatibility mode)."""
    from stochqn_b200.distributed import shard_bounds

    offset, n_local = shard_bounds(n, rank, world)
    hostcb_path = os.path.join(ROOT, "stochqn_b200", "lib", "libhostcb_f64.so")
    if not os.path.exists(hostcb_path):
        return None
    hostcb = C.CDLL(hostcb_path)
    hostcb.host_rosenbrock_grad.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_longlong, C.c_double, C.c_double]
    hostcb.host_rosenbrock_x0.argtypes = [C.c_void_p, C.c_longlong, C.c_longlong]
    host_threads = max(1, (os.cpu_count() or 1) // max(world, 1))
    hostcb.host_set_threads(host_threads)
    steps = min(args.steps, args.e2e_steps)
    warmup = MEM + 2
    xh = torch.empty(n_local, dtype=torch.float64).pin_memory()
    gh = torch.empty(n_local, dtype=torch.float64).pin_memory()
    hostcb.host_rosenbrock_x0(xh.data_ptr(), n_local, offset)
    ws = lib.initialize_oLBFGS(n_local, MEM, 0.0, 0.0, MIN_CURV, 1, 1)
    if not ws:
        return None
    if comm is not None:
        lib.stochqn_b200_set_comm(ws, comm, n)
    req, task, info = C.c_void_p(), C.c_int(), C.c_int()
    xp, gp = xh.data_ptr(), gh.data_ptr()
    gloo = dist.new_group(backend="gloo") if dist is not None else None

    def serve():
        hl = hr = 0.0
        if world > 1:
            mine = torch.tensor([float(xh[0]), float(xh[-1])], dtype=torch.float64)
            parts = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
            dist.all_gather(parts, mine, group=gloo)
            hl = float(parts[rank - 1][1]) if rank > 0 else 0.0
            hr = float(parts[rank + 1][0]) if rank < world - 1 else 0.0
        hostcb.host_rosenbrock_grad(req.value, gp, n_local, offset, n, hl, hr)

    split = {"callback_s": 0.0, "step_call_s": 0.0, "pair_call_s": 0.0}

    def iteration():
        t_a = time.perf_counter()
        serve()
        t_b = time.perf_counter()
        lib.run_oLBFGS(STEP, xp, gp, C.byref(req), C.byref(task), ws, C.byref(info))
        t_c = time.perf_counter()
        split["callback_s"] += t_b - t_a
        split["step_call_s"] += t_c - t_b
        if task.value == 102:
            serve()
            t_d = time.perf_counter()
            lib.run_oLBFGS(STEP, xp, gp, C.byref(req), C.byref(task), ws, C.byref(info))
            t_e = time.perf_counter()
            split["callback_s"] += t_d - t_c
            split["pair_call_s"] += t_e - t_d

    lib.run_oLBFGS(STEP, xp, gp, C.byref(req), C.byref(task), ws, C.byref(info))
    for _ in range(warmup):
        iteration()
    assert int(ws.contents.bfgs_memory.contents.mem_used) == MEM
    for k in split:
        split[k] = 0.0
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        iteration()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    opts = {"grad_writeback": int(lib.stochqn_b200_get_option(ws, 1)), "trust_x_mirror": int(lib.stochqn_b200_get_option(ws, 2))} \
        if hasattr(lib, "stochqn_b200_get_option") else {}
    lib.dealloc_oLBFGS(ws)
    vec = n_local * 8
    # per iteration and rank: step call uploads grad (and x unless the device mirror is trusted), downloads x (and grad when
    # write-back is on); pair call uploads grad
    h2d = (2 + (0 if opts.get("trust_x_mirror", 0) else 1)) * vec * world
    d2h = (1 + (1 if opts.get("grad_writeback", 1) else 0)) * vec * world
    return {"value": steps / dt, "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * dt / steps, "options": opts,
            "ms_split_rank0": {"host_gradient_callbacks": 1e3 * split["callback_s"] / steps,
                               "run_oLBFGS_step_call (H2D; K1-K3; D2H)": 1e3 * split["step_call_s"] / steps,
                               "run_oLBFGS_pair_call (H2D grad; K4)": 1e3 * split["pair_call_s"] / steps},
            "path": "run_oLBFGS with host pointers (pinned), host C+OpenMP gradient callback, %d host threads per rank" % host_threads}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=12)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=N_DEFAULT)
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--ref-budget", type=float, default=150.0, help="reference arm: seconds of CPU work the sample may take")
    ap.add_argument("--secondary-budget", type=float, default=60.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-pointer end-to-end leg (profiling runs)")
    ap.add_argument("--no-secondary", action="store_true", help="skip BASELINE configs 1, 2, 3, 5")
    ap.add_argument("--no-checks", action="store_true", help="skip the sharded-parity pre-check (profiling runs)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return main_reference(args)
    return main_b200(args)


if __name__ == "__main__":
    sys.exit(main())
