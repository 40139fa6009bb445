#!/bin/bash
for m in in out 0; do echo "LGW_DIRECT=$m"; LGW_DIRECT=$m python tools/exp_e2e.py 2>&1 | grep "slices"; done
for m in in out 0; do echo "TRACE LGW_DIRECT=$m"; SL=8 LGW_TRACE=1 LGW_DIRECT=$m python tools/exp_e2e.py 2>&1 | grep "lgw trace" | tail -9; done
timeout 900 python -m pytest tests/test_sse_gpu.py -m gpu -x -q -k "direct" 2>&1 | tail -4
