#!/bin/bash
# Final evidence of a round: full GPU suite, probes, the bench line, the launch list and the full capture of K1 / K3 / K4.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest --timeout=120 tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
(timeout 100 python tools/probe_fit.py cfg1; timeout 100 python tools/probe_fit.py cfg2) > gpurun_out/probe_fit.jsonl 2>&1
timeout 200 python tools/probe_cfg3.py > gpurun_out/probe_cfg3.json 2>&1
timeout 300 python tools/probe_gemm2.py > gpurun_out/probe_gemm2.jsonl 2>&1
timeout 700 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; echo "bench ref rc=$?"
PROF="python bench.py --steps 2 --warmup 12 --no-e2e --no-cpu-baseline --no-secondary --no-checks"
timeout 300 $PROF > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err; rc=$?; echo "prof plain rc=$rc"
if [ $rc -eq 0 ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $PROF > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
fi
timeout 200 python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke rc=$?"
