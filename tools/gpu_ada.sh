#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 500 python -m pytest --timeout=60 tests/test_gpu_device_loop.py tests/test_gpu_multinomial.py tests/test_gpu_logistic_estimator.py tests/test_gpu_guided.py -x -q > gpurun_out/pytest_ada.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_ada.log
timeout 200 python tools/probe_cfg3.py
timeout 300 python tools/bench_configs.py cfg3d --steps 1000 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['config'], d.get('loop'), 'steps/s %.0f'%d['steps_per_s'], 'us/step %.1f'%(d['ms_per_step']*1e3), 'launches/step %.2f'%d['launches_per_step'], 'loss %.5f'%d['loss_after'], d['infos'], d.get('device_loop_steps'))
    else: print(l.rstrip()[:300])
"
