#!/bin/bash
# Build the reference's own sources, read in place from $CLOUDAAE_REFERENCE (default /root/reference),
# into oracle/_ref/libcloudaae_ref.so.  TEST INFRASTRUCTURE ONLY.  No reference source is copied.
#   - tf_ops/nn_distance/tf_nndistance.cpp  : g++ -O2 (as tf_nndistance_compile.sh:3) against the
#                                             header stand-in in oracle/ref_shim/
#   - tf_ops/nn_distance/tf_nndistance_g.cu : nvcc -O2 -DGOOGLE_CUDA=1 (as :1), arch sm_100a
#   - tf_ops/sampling/tf_sampling_g.cu      : nvcc -O2 (as tf_sampling_compile.sh:2), arch sm_100a
# The GPU box has no /root/reference; it uses the prebuilt .so that travels with the snapshot.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${CLOUDAAE_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/tf_ops" ]; then
  echo "build_ref: $REF not present; keeping prebuilt $OUT (if any)"; exit 0
fi
mkdir -p "$OUT"
ARCH="-gencode arch=compute_100a,code=sm_100a"
nvcc -O2 -DGOOGLE_CUDA=1 $ARCH -x cu -Xcompiler -fPIC -c "$REF/tf_ops/nn_distance/tf_nndistance_g.cu" -o "$OUT/tf_nndistance_g.o"
nvcc -O2 -DGOOGLE_CUDA=1 $ARCH -x cu -Xcompiler -fPIC -c "$REF/tf_ops/sampling/tf_sampling_g.cu" -o "$OUT/tf_sampling_g.o"
/usr/bin/g++ -std=c++11 -O2 -ffp-contract=off -fPIC -I "$HERE/ref_shim" -c "$REF/tf_ops/nn_distance/tf_nndistance.cpp" -o "$OUT/tf_nndistance.o"
/usr/bin/g++ -std=c++11 -O2 -fPIC -I "$HERE/ref_shim" -c "$HERE/ref_shim/ref_driver.cpp" -o "$OUT/ref_driver.o"
nvcc -shared $ARCH -o "$OUT/libcloudaae_ref.so" "$OUT/tf_nndistance_g.o" "$OUT/tf_sampling_g.o" "$OUT/tf_nndistance.o" "$OUT/ref_driver.o"
rm -f "$OUT"/*.o
echo "build_ref: built $OUT/libcloudaae_ref.so"
