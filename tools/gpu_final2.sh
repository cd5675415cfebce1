#!/bin/bash
# Final 1-GPU visit of the round: the in-process exchange tests first, then the rest of the GPU suite.
set -u
mkdir -p gpurun_out
timeout 700 python -m pytest tests/test_gpu_zz_inprocess.py -q > gpurun_out/pytest_inproc.log 2>&1; echo "inproc rc=$?"
tail -12 gpurun_out/pytest_inproc.log
timeout 600 python -m pytest tests -m gpu -x -q --ignore=tests/test_gpu_zz_inprocess.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
