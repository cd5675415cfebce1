"""Worker of tests/test_gpu_multi.py::test_row_sharded_multinomial: adaQN + multinomial callbacks with the batch rows
sharded over 2 GPUs, both combination modes, against the same optimisation run on ONE GPU over the union of the rows.
argv: out_json"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tools"))

import bench_configs as BC          # noqa: E402


def main():
    out = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    res = {}
    xs = {}
    for mode in ("allreduce", "zero1", "p2p"):
        r, x = BC.run_multinomial_sharded("t", np.float64, 64, 40, 32, 512, 45, 5, 0.9, 1e-2, mode=mode, warm_cycles=2, return_x=True, quiet=True)
        res[mode] = dict(x_norm=r["x_norm"], tasks=r["tasks"], infos=r["infos"], mem_used=r["mem_used"])
        xs[mode] = x
    # fp32, tensor-core shapes: the reduce-scatter fused into the GEMM epilogue (peer-memory stores) against ncclReduceScatter
    fused = {}
    for mode in ("zero1", "fused", "p2p"):
        r, x = BC.run_multinomial_sharded("t32", np.float32, 1024, 512, 256, 2048, 24, 4, 0.9, 1e-3, mode=mode, warm_cycles=1, return_x=True, quiet=True)
        fused[mode] = dict(x=x, tasks=r["tasks"], infos=r["infos"])
    fused_err = float(np.max(np.abs(fused["zero1"]["x"] - fused["fused"]["x"])) / max(np.max(np.abs(fused["zero1"]["x"])), 1e-30))
    fused_same_tasks = fused["zero1"]["tasks"] == fused["fused"]["tasks"] and fused["zero1"]["infos"] == fused["fused"]["infos"]
    fused_moved = float(np.max(np.abs(fused["fused"]["x"])))
    # push all-gather + pull reduce-scatter over peer memory against the NCCL collectives (same arithmetic up to the order of the sum over ranks)
    p2p32_err = float(np.max(np.abs(fused["zero1"]["x"] - fused["p2p"]["x"])) / max(np.max(np.abs(fused["zero1"]["x"])), 1e-30))
    p2p32_same_tasks = fused["zero1"]["tasks"] == fused["p2p"]["tasks"] and fused["zero1"]["infos"] == fused["p2p"]["infos"]
    p2p_err = float(np.max(np.abs(xs["zero1"] - xs["p2p"])) / np.max(np.abs(xs["zero1"])))
    # the two modes run the same arithmetic up to summation order
    err = float(np.max(np.abs(xs["allreduce"] - xs["zero1"])) / np.max(np.abs(xs["allreduce"])))
    # every rank must hold the same x
    t = torch.tensor(xs["zero1"], device="cuda")
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    same = all(bool(torch.equal(parts[0], p)) for p in parts)
    if rank == 0:
        json.dump(dict(res=res, modes_rel_err=err, ranks_identical=same, moved=float(np.max(np.abs(xs["zero1"]))),
                       fused_rel_err=fused_err, fused_same_tasks=fused_same_tasks, fused_moved=fused_moved,
                       p2p_rel_err=p2p_err, p2p32_rel_err=p2p32_err, p2p32_same_tasks=p2p32_same_tasks), open(out, "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
