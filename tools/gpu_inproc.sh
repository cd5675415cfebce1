#!/bin/bash
# One short GPU visit: the in-process multi-rank tests (one GPU).
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zz_inprocess.py -q --timeout 580 > gpurun_out/pytest_inproc.log 2>&1; echo "inproc rc=$?"
tail -25 gpurun_out/pytest_inproc.log
