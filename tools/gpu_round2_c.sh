#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sse_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c3_g.json 2> gpurun_out/r02_bench_c3_g.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_c3_g.json').read())
print({k:d[k] for k in ('value','ms_per_step','kernel_ms','segments')}, d['roofline']['frac'], d['e2e']['ms_per_step'])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_relay2 -s 3 -c 1 -o gpurun_out/r02_relay2_g -f python tools/exp_one.py > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_commit2 -s 3 -c 1 -o gpurun_out/r02_commit_g -f python tools/exp_one.py > /dev/null 2>&1
ls -la gpurun_out/*_g.ncu-rep
