#!/bin/bash
# usage: bash tools/gpu_lib_ab.sh build/variants/libX.so [repeats]  — step time with the in-tree library vs a variant library,
# alternating on the same box (box-to-box variance is ~0.5 %, larger than most single-kernel effects)
set -u
V=$1; R=${2:-3}
L=cloudaae_b200/lib/libcloudaae_b200.so
cp $L /tmp/lib_head.so
for i in $(seq $R); do
  cp /tmp/lib_head.so $L; echo "head    $(timeout 60 python tools/ab_pipeline.py 1 2>&1 | tail -1 | cut -c1-30)"
  cp $V $L;               echo "variant $(timeout 60 python tools/ab_pipeline.py 1 2>&1 | tail -1 | cut -c1-30)"
done
cp /tmp/lib_head.so $L
