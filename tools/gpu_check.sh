#!/bin/bash
# One gpurun call: SSE parity tests, the perf probe, a short bench line.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sse_gpu.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/check_pytest.log
timeout 300 python tools/exp_relay_variants.py 2>&1 | tee gpurun_out/check_variants.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/check_bench.json 2> gpurun_out/check_bench.err; tail -c 3000 gpurun_out/check_bench.json; tail -5 gpurun_out/check_bench.err
