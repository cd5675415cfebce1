#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multinomial.py tests/test_gpu_logistic_estimator.py tests/test_gpu_guided.py -x -q > gpurun_out/pytest_mn.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_mn.log
for v in 1 0; do
echo "== MN_SMALL=$v"
STOCHQN_B200_MN_SMALL=$v timeout 300 python tools/bench_configs.py cfg3 --steps 1000 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['config'], 'steps/s %.0f'%d['steps_per_s'], 'us/step %.1f'%(d['ms_per_step']*1e3), 'launches/step %.2f'%d['launches_per_step'], 'loss %.5f'%d['loss_after'], d['infos'])
    else: print(l.rstrip()[:300])
"
done
timeout 120 python tools/probe_multinomial.py cfg3 2>&1 | tail -5
