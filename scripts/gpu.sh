#!/bin/bash
# Build in-tree, then run a command on the GPU box: scripts/gpu.sh [--gpus N] <timeout_s> '<command>'
set -e
cd "$(dirname "$0")/.."
python fem_2d_b200/build.py >/dev/null
G=""
if [ "$1" == "--gpus" ]; then G="--gpus $2"; shift 2; fi
T=$1; shift
exec /usr/local/graft/bin/gpurun $G --timeout "$T" -- "$@"
