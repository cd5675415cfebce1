#!/usr/bin/env python
"""bench.py - optimizer steps/s of free-mode oLBFGS on the chained Rosenbrock function.

Workload = BASELINE.json config 4: fp64, n = 2^27 parameters, mem_size 10,
hess_init 0 (gamma scaling), y_reg 0, min_curvature 1e-4, check_nan 1, step 1e-4,
x0[i] = 0.95 + 1e-4*((uint32)(i*2654435761) mod 1000).  One "step" = one oLBFGS iteration =
serve calc_grad -> run_oLBFGS (take the step) -> serve calc_grad_same_batch -> run_oLBFGS (pair).

    python bench.py [--gpus N] [--steps K] [--warmup W]            this repo (CUDA, sm_100a)
    python bench.py --impl reference [...]                          the reference's own CPU build (oracle/_ref)

N > 1: launched by torchrun, one rank per GPU; the parameter vector (total length fixed at n: strong
scaling) shards by contiguous blocks, dot partials go through one small NCCL all-reduce per phase.

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (x, gradients, pairs all in HBM,
gradient requests served by the bundled device callback).  `e2e` = the same metric through the drop-in
C ABI with HOST buffers: x and grad live in pinned host memory, the gradient is evaluated by a host
(C + OpenMP) callback as a user of the reference would, and every call stages host<->device copies.
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
METRIC = "optimizer steps/s at n=2^27, m=10 (oLBFGS, fp64, chained Rosenbrock)"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


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
    streaming pass, cost linear in n)."""
    if not os.path.exists(_harness_path()):
        return None
    cores = os.cpu_count() or 1
    # calibrate on a short run
    cal_n = 2 ** 22
    best = None
    for th in sorted({1, cores}):
        r = _run_harness(cal_n, 11, 3, th)
        if best is None or r["steps_per_s"] > best[1]["steps_per_s"]:
            best = (th, r)
    th, r = best
    per_elem_step = 1.0 / (r["steps_per_s"] * cal_n)                 # seconds per element per iteration
    n_s = n_target
    while n_s > 2 ** 16 and per_elem_step * n_s * (warmup + steps) * 1.3 > budget_s:
        n_s //= 2
    rr = _run_harness(n_s, warmup, steps, th)
    scale = n_s / float(n_target)
    return dict(value=rr["steps_per_s"] * scale, opt_only=rr["opt_steps_per_s"] * scale, cores=th, host_cores=cores, n_sample=n_s,
                raw_steps_per_s=rr["steps_per_s"], seconds=rr["seconds"], steps=steps, warmup=warmup,
                info_events=rr["info_events"], x_norm=rr["x_norm"])


def main_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    t0 = time.time()
    r = cpu_reference_run(args.steps, args.warmup, budget_s=150.0)
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/rosen_harness_f64 missing (reference not built)"}))
        return 0
    sample = ("reference C library (unmodified src/stochqn.c, gcc -O2 -fopenmp, OpenBLAS) on the same oLBFGS/Rosenbrock "
              "workload at n=%d (1/%d of 2^27), %d warm-up + %d timed iterations, %d thread(s) of %d host cores; "
              "steps/s scaled by n_sample/2^27 (all passes are linear in n)"
              % (r["n_sample"], N_DEFAULT // r["n_sample"], r["warmup"], r["steps"], r["cores"], r["host_cores"]))
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / r["value"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "BASELINE config 4: free-mode oLBFGS, chained Rosenbrock, n=2^27, mem_size=10, fp64",
                   "n": N_DEFAULT, "mem_size": MEM, "step": STEP, "min_curvature": MIN_CURV, "check_nan": 1},
        "cpu_baseline": {"value": r["value"], "unit": "steps/s", "cores": r["cores"], "kind": "reference", "sample": sample},
        "e2e": {"value": r["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.time() - t0,
    }
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
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
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
# this repo's arm
# ------------------------------------------------------------------------------------------------
def main_b200(args):
    import numpy as np
    import torch

    from stochqn_b200 import _lib
    from stochqn_b200.distributed import init_comm, shard_bounds

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
    offset, n_local = shard_bounds(n, rank, world)
    comm = init_comm(abi, rank, world) if world > 1 else None
    stream = torch.cuda.current_stream().cuda_stream      # 0 = legacy default stream

    x = torch.empty(n_local, device="cuda", dtype=torch.float64)
    g = torch.empty(n_local, device="cuda", dtype=torch.float64)
    halo = torch.zeros(2, device="cuda", dtype=torch.float64)
    scratch = torch.zeros(2 * max(world, 1), device="cuda", dtype=torch.float64)
    lib.stochqn_b200_rosenbrock_x0(x.data_ptr(), n_local, offset, stream)
    ws = lib.initialize_oLBFGS(n_local, MEM, 0.0, 0.0, MIN_CURV, 1, 1)
    if not ws:
        raise SystemExit("initialize_oLBFGS failed: " + _lib.last_error(abi))
    if comm is not None:
        assert lib.stochqn_b200_set_comm(ws, comm, n) == 0
    req, task, info = C.c_void_p(), C.c_int(), C.c_int()
    xp, gp, hp, sp = x.data_ptr(), g.data_ptr(), halo.data_ptr(), scratch.data_ptr()
    events = {"info": 0}

    def serve_gradient():
        if world > 1:       # halo exchange fused into the gradient kernel (peer memory), one launch
            lib.stochqn_b200_rosenbrock_grad_sharded(req.value, gp, n_local, offset, n, rank, world, comm, hp, sp, stream)
        else:
            lib.stochqn_b200_rosenbrock_grad(req.value, gp, n_local, offset, n, hp, stream)

    def iteration():
        serve_gradient()                                                                          # calc_grad
        lib.run_oLBFGS(STEP, xp, gp, C.byref(req), C.byref(task), ws, C.byref(info))              # step
        events["info"] += info.value != 200
        if task.value == 102:
            serve_gradient()                                                                      # calc_grad_same_batch
            lib.run_oLBFGS(STEP, xp, gp, C.byref(req), C.byref(task), ws, C.byref(info))          # pair
            events["info"] += info.value != 200

    def fence():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    lib.run_oLBFGS(STEP, xp, gp, C.byref(req), C.byref(task), ws, C.byref(info))                  # section 0
    sampler = ClockSampler(local_rank) if rank == 0 else None      # samples the warm-up and the timed region (both under load)
    for _ in range(args.warmup):
        iteration()
    lib.stochqn_b200_set_option(ws, _lib.OPT_PROFILE, 1)
    fence()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        iteration()
    e1.record()
    fence()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = args.steps / (ms * 1e-3)

    st = {k: _lib.get_stat(abi, ws, v) for k, v in dict(k1=1, k1n=2, k3=3, k3n=4, k4=5, k4n=6).items()}
    lib.stochqn_b200_set_option(ws, _lib.OPT_PROFILE, 0)
    used = int(ws.contents.bfgs_memory.contents.mem_used)
    x_norm2 = torch.sum(x * x)
    if dist is not None:
        dist.all_reduce(x_norm2)
    x_norm = float(torch.sqrt(x_norm2).item())
    lib.dealloc_oLBFGS(ws)
    del g

    # ---- roofline of the dominant kernel (K3: fused combine + update), per launch ------------------
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
    k3_bytes = (2 * used + 4) * vec_bytes              # read g, S, Y, x; write x, s_new  (SURVEY 8(d)); the grad write-back is not counted
    k1_bytes = (2 * used + 2) * vec_bytes              # read g, S, Y; write grad_prev
    k4_bytes = 4 * vec_bytes
    traffic, traffic_note = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        # ncu --set full was taken at n = 2^24 (40 replays of a 20 GiB working set are not practical at 2^27); every byte
        # of K3's traffic is streaming, so the per-launch figure scales with the vector length
        traffic = float(tj["k3_dram_bytes_per_launch"]) * n_local / float(tj["n_local"])
        traffic_note = "dram__bytes_read+write of one K3 launch captured at n_local=%d, scaled by n_local/%d (%s)" % (
            int(tj["n_local"]), int(tj["n_local"]), tj.get("source", "profiles/"))
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "k3_combine (fused combine + x update + new s)", "achieved": k3_bytes / k3_ms / 1e6,
                "peak": peak, "unit": "GB/s", "frac": k3_bytes / k3_ms / 1e6 / peak, "traffic": traffic, "traffic_note": traffic_note,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                "per_launch_bytes": k3_bytes, "avg_launch_ms": k3_ms, "launches_timed": int(st["k3n"]),
                "other_kernels": {
                    "k1_dots": {"achieved": k1_bytes / k1_ms / 1e6, "frac": k1_bytes / k1_ms / 1e6 / peak, "avg_launch_ms": k1_ms},
                    "k4_pair": {"achieved": k4_bytes / k4_ms / 1e6, "frac": k4_bytes / k4_ms / 1e6 / peak, "avg_launch_ms": k4_ms}},
                "step": {"algorithmic_bytes": (4 * used + 14) * vec_bytes * world, "achieved": (4 * used + 14) * vec_bytes * world / (ms / args.steps) / 1e6,
                         "frac_of_aggregate_peak": (4 * used + 14) * vec_bytes * world / (ms / args.steps) / 1e6 / (peak * world)}}

    # ---- end-to-end through the drop-in C ABI with HOST buffers --------------------------------------
    e2e = None if args.no_e2e else run_e2e(args, lib, abi, rank, world, local_rank, n, offset, n_local, comm, dist, torch, np)

    # ---- CPU baseline (rank 0, N = 1 only) -------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(8, 12, budget_s=25.0)
        if r is not None:
            cpu = {"value": r["value"], "unit": "steps/s", "cores": r["cores"], "kind": "reference",
                   "sample": "reference C library (oracle/_ref, unmodified src/stochqn.c + OpenBLAS) on the same workload at n=%d "
                             "(1/%d of 2^27), 12 warm-up + 8 timed iterations, %d thread(s) of %d host cores; scaled by n_sample/2^27"
                             % (r["n_sample"], N_DEFAULT // r["n_sample"], r["cores"], r["host_cores"]),
                   "optimizer_only_value": r["opt_only"]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "BASELINE config 4: free-mode oLBFGS, chained Rosenbrock, n=2^27, mem_size=10, fp64, sharded by contiguous blocks",
                       "n": n, "n_per_gpu": n_local, "mem_size": MEM, "step": STEP, "min_curvature": MIN_CURV, "check_nan": 1,
                       "grad_writeback": 1, "l2": "inputs larger than L2 (%.1f GiB streamed per step per GPU)" % ((4 * used + 14) * vec_bytes / 2 ** 30),
                       "callbacks": "bundled device Rosenbrock gradient (+1 all-reduce halo when sharded)"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "check": {"x_norm": x_norm, "info_events": events["info"], "mem_used": used},
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        if comm is not None:
            lib.stochqn_b200_comm_destroy(comm)
        dist.destroy_process_group()
    return 0


def run_e2e(args, lib, abi, rank, world, local_rank, n, offset, n_local, comm, dist, torch, np):
    """Same metric through run_oLBFGS with HOST pointers: x / grad in pinned host memory, gradient by a host
    (C + OpenMP) callback, H2D / D2H staging inside every call (the library's compatibility mode)."""
    from stochqn_b200 import _lib

    hostcb_path = os.path.join(ROOT, "stochqn_b200", "lib", "libhostcb_f64.so")
    if not os.path.exists(hostcb_path):
        return None
    hostcb = C.CDLL(hostcb_path)
    hostcb.host_rosenbrock_grad.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_longlong, C.c_double, C.c_double]
    hostcb.host_rosenbrock_x0.argtypes = [C.c_void_p, C.c_longlong, C.c_longlong]
    host_threads = max(1, (os.cpu_count() or 1) // max(world, 1))
    hostcb.host_set_threads(host_threads)
    steps = min(args.steps, args.e2e_steps)
    warmup = max(3, min(args.warmup, 12))
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
    edge = torch.zeros(2 * world, dtype=torch.float64)
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
    lib.dealloc_oLBFGS(ws)
    del edge
    vec = n_local * 8
    # per iteration and rank: step call uploads x and grad, downloads x and grad (write-back on); pair call uploads grad
    return {"value": steps / dt, "unit": "steps/s", "h2d_bytes_per_step": 3 * vec * world, "d2h_bytes_per_step": 2 * vec * world,
            "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * dt / steps,
            "ms_split_rank0": {"host_gradient_callbacks": 1e3 * split["callback_s"] / steps,
                               "run_oLBFGS_step_call (H2D x, grad; K1-K3; D2H x, grad)": 1e3 * split["step_call_s"] / steps,
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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-pointer end-to-end leg (profiling runs)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return main_reference(args)
    return main_b200(args)


if __name__ == "__main__":
    sys.exit(main())
