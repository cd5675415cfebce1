#!/bin/bash
# One GPU-box visit: parity tests, the bench line, the ncu launch list and one full capture of K1/K3/K4.
# Every ncu pass runs only after the same command exited 0 without ncu.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err; echo "bench ref rc=$?"
PROF="python bench.py --steps 2 --warmup 12 --no-e2e --no-cpu-baseline --no-secondary --no-checks"
timeout 300 $PROF > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err; rc=$?; echo "prof plain rc=$rc"
if [ $rc -eq 0 ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $PROF > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k1_dots|k3_combine|k4_pair' --launch-skip 36 --launch-count 3 -o gpurun_out/full_k1k3k4 -f $PROF > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
  ncu -i gpurun_out/full_k1k3k4.ncu-rep --page raw --csv > gpurun_out/full_k1k3k4_raw.csv 2>/dev/null
fi
