#!/bin/bash
# One short gpurun call: SSE + transcript parity tests and a bench line without the CPU baseline.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sse_gpu.py tests/test_transcript_gpu.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/quick_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err; tail -3 gpurun_out/quick_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/quick_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','kernel_ms')}, d['roofline']['frac'], d['e2e']['ms_per_step'], d.get('transcript_tap'))
P
