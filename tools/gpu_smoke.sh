#!/bin/bash
# Shortest GPU visit: smoke() + the direct CUDA-vs-reference-library parity file.
set -u
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 150 python -m pytest tests/test_gpu_vs_reflib.py -q -x 2>&1 | tail -2
