#!/bin/bash
# 2-GPU visit: the bench line at N = 2 (sharded / row-sharded parity incl. the p2p mode, cfg5s zero1 vs p2p with phases) + the 2-GPU row-shard test
set -u
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e > gpurun_out/bench_n2_p2p.json 2> gpurun_out/bench_n2_p2p.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_n2_p2p.json").read().strip().splitlines()[-1])
    s = d.get("secondary", {})
    print("value", d.get("value"), "sharded_parity", d.get("check", {}).get("sharded_parity", {}).get("ok"))
    print(json.dumps({k: s.get(k) for k in ("rowsharded_parity", "cfg5s_modes", "cfg5s_phase_ms_by_mode", "wall_s")}, indent=0)[:2500])
except Exception as e:
    print("no line:", e)
PY
tail -5 gpurun_out/bench_n2_p2p.err
timeout 400 python -m pytest tests/test_gpu_multi.py -q -x -k row_sharded > gpurun_out/pytest_multi_rowshard.log 2>&1; echo "pytest rowshard rc=$?"
tail -15 gpurun_out/pytest_multi_rowshard.log
