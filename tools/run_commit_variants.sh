#!/bin/bash
mkdir -p gpurun_out
for lib in "" $(ls llmapigateway_b200/_native/variants/*.so); do
  name=$(basename "${lib:-default}" .so)
  echo "=== $name $(LGW_NATIVE_LIB=$lib timeout 200 python tools/exp_commit.py 2>&1 | tail -1)"
done | tee gpurun_out/commit_variants.log
