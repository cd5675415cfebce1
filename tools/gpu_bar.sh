#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest --timeout=60 tests/test_gpu_device_loop.py tests/test_gpu_logistic_estimator.py tests/test_gpu_guided.py -x -q > gpurun_out/pytest_bar.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_bar.log
timeout 120 python tools/probe_fit.py cfg1 | cut -c1-600
timeout 120 python tools/probe_cfg3.py | cut -c1-900
timeout 300 python tools/bench_configs.py cfg1d cfg2d cfg3d --steps 2000 --rows-cfg2 200000 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['config'], d['loop'][:30], 'steps/s %.0f'%d['steps_per_s'], 'us/step %.1f'%(d['ms_per_step']*1e3), 'launches/step %.2f'%d['launches_per_step'], d['infos'])
    else: print(l.rstrip()[:300])
"
