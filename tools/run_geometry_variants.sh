#!/bin/bash
# Kernel-geometry experiment (one gpurun call): the C3 step and its variants with the bulk kernel built for other block shapes
# (warps per block x tile bytes x ring depth).  Each library is a full build of the same sources (make OUT=... EXTRA="-DR2_...").
mkdir -p gpurun_out
for lib in "" $(ls llmapigateway_b200/_native/variants/*.so); do
  name=$(basename "${lib:-default}" .so)
  echo "=== $name"
  LGW_NATIVE_LIB=$lib timeout 200 python tools/exp_relay_variants.py 2>&1 | grep -v "templates:" | cut -c1-330
done | tee gpurun_out/geometry_variants.log
