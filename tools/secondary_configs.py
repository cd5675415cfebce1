"""bench.py's `secondary` block: the other BASELINE.json configurations, bounded to about a minute.

N = 1: configs 1, 2, 3, 5 device-resident (tools/bench_configs.py), each with
  * its algorithmic bytes / flops per step (SURVEY.md section 8(d) formulas, written out below),
  * the fraction of the roofline that step time corresponds to (measured HBM copy bandwidth; tensor work at half the
    measured dense bf16 rate = the tf32 rate), and
  * the reference C library (oracle/_ref) + NumPy/BLAS callbacks timed on the host cores on rows of the same data
    (tools/cpu_configs.py; the only place this module touches oracle/).
N > 1: config 5 with the batch rows sharded over the ranks (reduce-scatter -> sharded adaQN step -> all-gather), per-phase
  device times, and a parity check of the row-sharded modes against the unsharded optimisation of the union of the rows.
"""
from __future__ import annotations

import os
import time

import numpy as np
import torch

import bench_configs as BC

MEM = 10


def _roof(bytes_per_step, flops_per_step, ms_per_step, peaks):
    hbm = float(peaks.get("hbm_gbs", 6650.0)) * 1e9
    tf32 = float(peaks.get("bf16_tflops", 1677.6)) * 1e12 / 2.0
    t_min = bytes_per_step / hbm + flops_per_step / tf32
    return {"algorithmic_bytes_per_step": bytes_per_step, "tensor_flops_per_step": flops_per_step,
            "roofline_ms_per_step": 1e3 * t_min, "frac": 1e3 * t_min / ms_per_step,
            "achieved_GBps": bytes_per_step / (ms_per_step * 1e-3) / 1e9,
            "peaks": "hbm %.0f GB/s (measured copy), tensor %.0f TFLOP/s tf32 (= measured dense bf16 / 2)" % (hbm / 1e9, tf32 / 1e12)}


def _cfg1(peaks, cpu):
    import cpu_configs as CC
    n, B = 1001, 1000
    cpu_fn = None
    if cpu:
        def cpu_fn(X, y):
            Xh, yh = X[:20000].cpu().numpy(), y[:20000].cpu().numpy()
            return CC.best_of_threads(CC.logistic_reference, kind="oLBFGS", X=Xh, y=yh, batch=B, steps=300, warm=15)
    out = None
    for native in ("batches", True, False):     # device-side loop, the library's host-driven loop, the caller's loop: report the best, keep all
        r = BC.run_logistic("cfg1", "oLBFGS", 100000, 1000, B, 2000, native=native, quiet=True, cpu_fn=cpu_fn if out is None else None)
        if out is None:
            out = r
            out["loops"] = {}
        out["loops"][r["loop"]] = r["steps_per_s"]
        if r["steps_per_s"] > out["steps_per_s"]:
            keep = {k: out[k] for k in ("loops", "cpu_reference") if k in out}
            out = dict(r, **keep)
    # per step: two gradients, each ONE sweep of the batch (B x n), + the optimizer's (4m + 10) n-vectors
    b = 2 * B * n * 8 + (4 * MEM + 10) * n * 8
    out["dominant_kernel"] = "kl_fit_logistic (one launch per run of mini-batches: gradient sweeps + step + pair) - latency-bound: 8 KB vectors"
    out["roofline"] = _roof(b, 0, out["ms_per_step"], peaks)
    return out


def _cfg2(peaks, cpu):
    import cpu_configs as CC
    n, B, L, big = 4097, 2000, 10, 20000
    cpu_fn = None
    if cpu:
        def cpu_fn(X, y):
            Xh, yh = X[:40000].cpu().numpy(), y[:40000].cpu().numpy()
            return CC.best_of_threads(CC.logistic_reference, kind="SQN", X=Xh, y=yh, batch=B, steps=50, warm=120, L=L, big=big, step=1e-2)
    out = BC.run_logistic("cfg2", "SQN", 1000000, 4096, B, 1000, L=L, big=big, native="batches", quiet=True, cpu_fn=cpu_fn, step=1e-2)
    # per step: one gradient sweep (B x n); per L steps one fused Hessian-vector sweep of the big batch and ~9 vectors of
    # pair work; optimizer (4m + 6) n-vectors
    b = B * n * 8 + big * n * 8 / L + (4 * MEM + 6 + 9.0 / L) * n * 8
    out["dominant_kernel"] = "logistic_fused (gradient: %d x %d sweep; Hessian-vector: %d x %d sweep every %d steps)" % (B, n, big, n, L)
    out["roofline"] = _roof(b, 0, out["ms_per_step"], peaks)
    return out


def _cfg3(peaks, cpu):
    import cpu_configs as CC
    d, K, B, L, k = 1836, 159, 50, 20, 100
    n = K * (d + 1)
    cpu_fn = None
    if cpu:
        def cpu_fn(X, lab, x0):
            return CC.best_of_threads(CC.multinomial_reference, dtype=np.float64, X=X.cpu().numpy(), lab=lab.cpu().numpy().astype(np.int64), K=K,
                                      batch=B, steps=40, warm=12 * L, L=L, fisher=k, use_grad_diff=0, max_incr=1.01, rms=0.0, step=1e-2, max_s=4.0,
                                      alpha=1e-1, wsum=True, x0=x0)
    out = None
    for native in (True, False):        # the request loop inside the library (two launches per ordinary step) and the caller's Python loop
        r = BC.run_multinomial("cfg3", np.float64, d, K, B, 6655, 1000, L, k, 0, 1.01, 0.0, 1e-2, quiet=True, cpu_fn=cpu_fn if out is None else None,
                               profile="bibtex", native=native)
        if out is None:
            out = r
            out["loops"] = {}
        out["loops"][r["loop"]] = r["steps_per_s"]
        if r["steps_per_s"] > out["steps_per_s"]:
            keep = {k2: out[k2] for k2 in ("loops", "cpu_reference") if k2 in out}
            out = dict(r, **keep)
    out["workload"] = ("BibTeX-shaped as in example/example_stochqn.ipynb: 1836 binary features (3.75 % dense), 159 classes, batch 50, summed loss, "
                       "reg_param 0.1, x0 ~ N(0,1), step 1e-2, AdaGrad, Fisher 100, L 20, max_incr 1.01")
    # per step: optimizer 4m + 10 = 50 n-vectors (Fisher ring write included), gradient reads W, alpha*W and writes G (3 vectors)
    # + the batch; per L steps the Fisher product 2k + 4 vectors
    b = (4 * MEM + 10 + 3 + (2 * k + 4) / float(L)) * n * 8 + B * d * 8
    out["dominant_kernel"] = "kl_ada (one-launch adaQN step) + mn_grad_small (one-launch gradient); n = %d: 2.3 MB vectors, latency-bound" % n
    out["roofline"] = _roof(b, 4.0 * B * d * K * 0, out["ms_per_step"], peaks)      # fp64 build: GEMMs on the CUDA cores, not counted
    return out


def _cfg5(peaks, cpu):
    import cpu_configs as CC
    d, K, B, L = 8192, 4096, 1024, 10
    n = K * (d + 1)
    out = BC.run_multinomial("cfg5", np.float32, d, K, B, 16384, 100, L, 0, 1, 0.0, 0.9, 1e-3, quiet=True, fixed_big=True)
    out["workload"] = "X ~ N(0,1)/sqrt(d), mean loss, reg 1e-3, x0 = 0, step 1e-3, RMSProp 0.9, grad-diff pairs on a fixed big batch (L = 10 batches)"
    # per step: tensor work 2 products of 2*B*d*K flop; optimizer 4m + 9 = 49 n-vectors, gradient reads W, alpha*W, writes G (3)
    b = (4 * MEM + 9 + 3) * n * 4 + 2 * B * d * 4
    out["dominant_kernel"] = "ka3_combine / ka1_dots (HBM) + gemm_tf32 (tcgen05)"
    out["roofline"] = _roof(b, 4.0 * B * d * K, out["ms_per_step"], peaks)
    if cpu:
        # bounded sample: the same shapes with classes / 8 (n / 8), so that the 10-pair memory fills within seconds; every pass
        # of the step is linear in the number of classes, the rate is scaled by 1/8
        Ks = K // 8
        g = torch.Generator(device="cuda").manual_seed(3)
        Xs = (torch.randn(4096, d, device="cuda", dtype=torch.float32, generator=g) / d ** 0.5).cpu().numpy()
        lab = np.random.default_rng(4).integers(0, Ks, 4096)
        r = CC.best_of_threads(CC.multinomial_reference, dtype=np.float32, X=Xs, lab=lab, K=Ks, batch=B, steps=10, warm=12 * L, L=L, fisher=0,
                               use_grad_diff=1, max_incr=0.0, rms=0.9, step=1e-3, thread_counts=[os.cpu_count() or 1], max_s=6.0, fixed_big=True)
        r["value"] /= 8.0
        r["sample"] += "; classes %d (1/8 of %d), steps/s scaled by 1/8" % (Ks, K)
        out["cpu_reference"] = r
    return out


def _rowsharded_parity(rank, world, dist):
    """Row-sharded modes against the UNSHARDED optimisation of the union of the rows (rank 0), fp64.  allreduce: replicated
    optimizer; zero1: ncclReduceScatter + sharded optimizer + ncclAllGather; p2p: the same with the library's own push
    all-gather and pull reduce-scatter over NVLink peer memory."""
    shape = dict(dtype=np.float64, d=64, K=40, batch_per_gpu=32, nrows_per_gpu=512, steps=45, L=5, rms=0.9, step=1e-2)
    res = {}
    xs = {}
    modes = ["allreduce", "zero1", "p2p"]
    for mode in list(modes):
        try:
            r, x = BC.run_multinomial_sharded("parity", mode=mode, warm_cycles=2, return_x=True, quiet=True, **shape)
        except Exception as e:                               # noqa: BLE001  (raised on every rank together, see bench_configs)
            if mode != "p2p":
                raise
            res[mode] = {"ok": False, "error": "%s: %s" % (type(e).__name__, e)}
            modes.remove(mode)
            continue
        xs[mode] = x
        res[mode] = {"tasks": r["tasks"], "infos": r["infos"]}
    flags = torch.ones(len(modes), device="cuda")
    if rank == 0:
        ru, xu = BC.run_multinomial_sharded("parity", mode="allreduce", warm_cycles=2, return_x=True, quiet=True, union_world=world, **shape)
        for i, mode in enumerate(modes):
            err = float(np.max(np.abs(xs[mode] - xu)) / np.max(np.abs(xu)))
            same = res[mode]["tasks"] == ru["tasks"] and res[mode]["infos"] == ru["infos"]
            res[mode] = {"x_rel_err_vs_unsharded": err, "sequences_equal_unsharded": bool(same), "ok": bool(same and err <= 1e-9)}
            flags[i] = 1.0 if res[mode]["ok"] else 0.0
        res["moved"] = float(np.max(np.abs(xu)))
    dist.broadcast(flags, 0)                                 # every rank must take the same decision about the p2p mode
    ok = {m: bool(flags[i].item() > 0) for i, m in enumerate(modes)}
    res["ok"] = bool(ok.get("allreduce", False) and ok.get("zero1", False))
    res["p2p_ok"] = bool(ok.get("p2p", False))
    return res


def run(rank, world, comm, dist, peaks, budget_s=60.0):
    t0 = time.time()
    out = {}
    torch.cuda.empty_cache()
    if world == 1:
        cpu = os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libstochqn_ref_f64.so"))
        for name, fn in (("cfg1", _cfg1), ("cfg3", _cfg3), ("cfg5", _cfg5), ("cfg2", _cfg2)):
            if time.time() - t0 > budget_s:
                out[name] = {"skipped": "secondary budget of %.0f s used up" % budget_s}
                continue
            try:
                r = fn(peaks, cpu)
                out[name] = {k: r[k] for k in ("workload", "optimizer", "loop", "loops", "dtype", "n", "batch", "steps", "ms_per_step", "steps_per_s", "mem_used",
                                               "launches_per_step", "infos", "dominant_kernel", "roofline", "cpu_reference", "loss_after", "loss_at_zero") if k in r}
                if "cpu_reference" in r:
                    out[name]["gpu_over_cpu"] = r["steps_per_s"] / r["cpu_reference"]["value"]
            except Exception as e:
                out[name] = {"error": "%s: %s" % (type(e).__name__, e)}
            torch.cuda.empty_cache()
    else:
        os.environ["CFG5S_PHASES"] = "1"
        try:
            out["rowsharded_parity"] = _rowsharded_parity(rank, world, dist)
            d, K, B = 8192, 4096, 1024
            n = K * (d + 1)
            keys = ("mode", "n_gpus", "dtype", "n", "batch_per_gpu", "global_batch", "steps", "ms_per_step", "steps_per_s", "samples_per_s", "mem_used",
                    "infos", "phase_ms_rank0", "roofline")
            runs = {}
            # zero1 = the NCCL collectives; p2p = the library's own push all-gather + pull reduce-scatter over NVLink peer
            # memory, timed only when its parity run above was green.  The faster one is the cfg5s line, both are kept.
            for mode in ["zero1"] + (["p2p"] if out["rowsharded_parity"].get("p2p_ok") else []):
                try:
                    r, _ = BC.run_multinomial_sharded("cfg5 row-sharded", np.float32, d, K, B, 16384, 60, 10, 0.9, 1e-3, mode=mode, warm_cycles=12, quiet=True, fixed_big=True)
                except Exception as e:                       # noqa: BLE001
                    if mode == "zero1":
                        raise
                    runs[mode] = {"error": "%s: %s" % (type(e).__name__, e)}
                    continue
                # per rank and step: the full tensor work of its rows; optimizer on 1/world of the vectors; the gradient's W read, G write,
                # alpha*W read on the full vector; reduce-scatter + all-gather move (world-1)/world of the n-vector each way
                b = (4 * MEM + 9) * n * 4 / world + 3 * n * 4 + 2 * B * d * 4
                r["roofline"] = _roof(b, 4.0 * B * d * K, r["ms_per_step"], peaks)
                r["roofline"]["nvlink_bytes_per_rank_per_step"] = 2.0 * n * 4 * (world - 1) / world
                runs[mode] = {k: r[k] for k in keys if k in r}
            best = min((m for m in runs if "ms_per_step" in runs[m]), key=lambda m: runs[m]["ms_per_step"])
            out["cfg5s"] = runs[best]
            out["cfg5s_modes"] = {m: (runs[m].get("ms_per_step") if "ms_per_step" in runs[m] else runs[m]) for m in runs}
            out["cfg5s_phase_ms_by_mode"] = {m: runs[m].get("phase_ms_rank0") for m in runs if "phase_ms_rank0" in runs[m]}
        except Exception as e:
            out["cfg5s"] = {"error": "%s: %s" % (type(e).__name__, e)}
    out["wall_s"] = time.time() - t0
    return out
