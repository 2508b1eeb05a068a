#!/bin/bash
# Rebuilds everything in-tree (the .so files travel with the snapshot), then hands the command to gpurun.
# usage: tools/gpu.sh [--timeout S] [--gpus N] -- 'command'
set -e
cd "$(dirname "$0")/.."
make -s all 2>&1 | grep -E "error|Error" && exit 1
exec /usr/local/graft/bin/gpurun "$@"
