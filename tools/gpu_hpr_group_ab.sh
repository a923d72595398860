#!/bin/bash
# A/B of the HPR LP group width (lanes per point): variant libraries built with -DHPR_GROUP=4|16 under build/variants/
set -u
L=cloudaae_b200/lib/libcloudaae_b200.so
cp $L /tmp/lib_w8.so
for w in 8 4 16 8 4; do
  if [ $w = 8 ]; then cp /tmp/lib_w8.so $L; else cp build/variants/libcloudaae_b200_w$w.so $L; fi
  t=$(timeout 120 python -m pytest tests/test_gpu_synthesis.py -x -q 2>&1 | tail -1)
  s=$(timeout 60 python tools/ab_pipeline.py 1 2>&1 | tail -1 | cut -c1-40)
  echo "group $w: tests [$t]  step [$s]"
done
cp /tmp/lib_w8.so $L
