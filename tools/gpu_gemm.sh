#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest --timeout=120 tests/test_gpu_multinomial.py -x -q > gpurun_out/pytest_gemm.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gemm.log
timeout 200 python tools/probe_multinomial.py cfg5 2>&1 | tail -3
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:gemm_tf32' --launch-skip 6 --launch-count 2 -o gpurun_out/full_gemm_2sm -f python tools/probe_multinomial.py cfg5 > gpurun_out/ncu_gemm.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/full_gemm_2sm.ncu-rep --page raw --csv > gpurun_out/full_gemm_2sm_raw.csv 2>/dev/null
python - <<'P'
import csv
rows=list(csv.reader(open('gpurun_out/full_gemm_2sm_raw.csv')))
hdr=rows[0]
want=['Kernel Name','gpu__time_duration.sum','sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_tensor.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','launch__cluster_size','launch__grid_size']
tens=[h for h in hdr if 'tensor' in h.lower() and 'pct' in h]
for r in rows[2:]:
    for w in want+tens[:8]:
        if w in hdr: print(w, r[hdr.index(w)][:80])
    print()
P
timeout 300 python tools/bench_configs.py cfg5 --steps 100 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['config'], 'steps/s %.0f'%d['steps_per_s'], 'us/step %.1f'%(d['ms_per_step']*1e3), 'launches/step %.2f'%d['launches_per_step'])
"
