#!/bin/bash
# One short GPU visit: the in-process collectives (one GPU) + the multinomial callbacks.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_inprocess_collectives.py -q -x --timeout 500 > gpurun_out/pytest_inproc.log 2>&1; echo "inproc rc=$?"
tail -25 gpurun_out/pytest_inproc.log
timeout 300 python -m pytest tests/test_gpu_multinomial.py -q -x > gpurun_out/pytest_mn.log 2>&1; echo "mn rc=$?"
tail -3 gpurun_out/pytest_mn.log
