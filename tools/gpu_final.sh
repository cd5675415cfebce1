#!/bin/bash
# Final evidence of a round: full GPU suite, probes, the bench line.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest --timeout=120 tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
(timeout 100 python tools/probe_fit.py cfg1; timeout 100 python tools/probe_fit.py cfg2) > gpurun_out/probe_fit.jsonl 2>&1; cat gpurun_out/probe_fit.jsonl | cut -c1-400
timeout 200 python tools/probe_cfg3.py > gpurun_out/probe_cfg3.json 2>&1; cut -c1-900 gpurun_out/probe_cfg3.json
timeout 700 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
