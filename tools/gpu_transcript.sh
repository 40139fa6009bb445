#!/bin/bash
# One gpurun call: transcript-tap parity tests on the GPU + a short bench line (carries the transcript_tap block).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_transcript_gpu.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/transcript_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/transcript_bench.json 2> gpurun_out/transcript_bench.err; tail -c 1500 gpurun_out/transcript_bench.json; tail -5 gpurun_out/transcript_bench.err
