#!/bin/bash
# Step-kernel variants (ree_split.cu hooks) timed with tools/kernel_bench.py.
# Usage under gpurun: bash tools/xf_variants.sh <tag> [n] ; CFGS="ws:variant ..." selects the list
TAG=${1:-x}; N=${2:-4000000}
OUT=gpurun_out/xfvar_$TAG.log; : > $OUT
for cfg in ${CFGS:-1:0 1:1 1:2 0:0}; do
  echo "REE_XF_WS=${cfg%%:*} REE_XF_VARIANT=${cfg##*:}" >> $OUT
  REE_XF_WS=${cfg%%:*} REE_XF_VARIANT=${cfg##*:} timeout 300 python tools/kernel_bench.py --n $N 2>&1 | grep "philox, transform only" >> $OUT
done
cat $OUT
