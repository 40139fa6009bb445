#!/bin/bash
mkdir -p gpurun_out
for lib in "" $(ls llmapigateway_b200/_native/variants/tx*.so); do
  name=$(basename "${lib:-default}" .so)
  echo "=== $name $(LGW_NATIVE_LIB=$lib timeout 120 python tools/exp_text.py 2>&1 | tail -1)"
  if [ -n "$lib" ]; then echo "    tests: $(LGW_NATIVE_LIB=$lib timeout 200 python -m pytest tests/test_transcript_gpu.py -m gpu -q 2>&1 | tail -1)"; fi
done | tee gpurun_out/text_variants.log
