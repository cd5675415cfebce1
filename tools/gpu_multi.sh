#!/bin/bash
# Multi-GPU visit: bash tools/gpu_multi.sh N  (under gpurun --gpus N)
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi_n$N.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/pytest_multi_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n=$N rc=$?"
grep '^{' gpurun_out/bench_n$N.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d.get('e2e',{}).get('value'))
print('check',json.dumps(d.get('check',{}).get('sharded_parity'))[:1500])
print('secondary',json.dumps(d.get('secondary'))[:1500])
"
tail -5 gpurun_out/bench_n$N.err
