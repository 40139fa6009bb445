#!/bin/bash
# Last GPU call of round 2 (tight budget): the full GPU parity suite on the final sources, then the headline bench line.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_gpu_tests.log 2>&1; tail -n 4 gpurun_out/r02b_gpu_tests.log
timeout 420 python bench.py > gpurun_out/r02b_bench_line.json 2> gpurun_out/r02b_bench_line.err; tail -c 300 gpurun_out/r02b_bench_line.err
python - <<'P'
import json
try:
    d = json.load(open('gpurun_out/r02b_bench_line.json'))
    print({k: d[k] for k in ('value', 'ms_per_step', 'kernel_ms')}, 'frac', d['roofline']['frac'], 'e2e ms', d['e2e']['ms_per_step'],
          'verdicts-only', d.get('e2e_verdicts_only'), 'tap', d.get('transcript_tap'), 'cpu', d.get('cpu_baseline', {}).get('value'))
except Exception as ex:
    print('no bench line:', ex)
P
