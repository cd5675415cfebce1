#!/bin/bash
# bash tools/gpu_bench_n.sh N   (under gpurun --gpus N): the bench line only
set -u
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n=$N rc=$?"
grep '^{' gpurun_out/bench_n$N.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value',d['value'],'ms',d['ms_per_step'],'spread',d.get('block_spread'),'e2e',d.get('e2e',{}).get('value'))
print('roofline', d['roofline']['frac'], d['roofline']['step'])
print('check',json.dumps(d.get('check',{}).get('sharded_parity'))[:600])
print('secondary',json.dumps(d.get('secondary'))[:1200])
"
tail -3 gpurun_out/bench_n$N.err
