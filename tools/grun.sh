#!/bin/bash
# Build in-tree (the .so files travel with the snapshot), then run a command on the GPU box:  tools/grun.sh [--gpus N] [--timeout S] -- 'command'
set -e
cd "$(dirname "$0")/.."
python -c "from stochqn_b200 import build; build.build()"
exec /usr/local/graft/bin/gpurun "$@"
