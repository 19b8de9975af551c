#!/bin/bash
# Rebuild libree_b200.so, then run the given command on a B200 box: tools/grun.sh <timeout_s> '<command>'
set -e
cd "$(dirname "$0")/.."
make -C rareeventestimation_b200/csrc 2>&1 | grep -v "^nvcc\|Nothing to be done\|Entering\|Leaving" || true
T=$1; shift
exec gpurun --timeout "$T" -- "$@"
