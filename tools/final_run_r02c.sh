#!/bin/bash
# GPU call after the ring change (R2_NBUF 3 -> 2): full parity suite, headline bench line, the step's variants.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_gpu_tests.log 2>&1; tail -n 3 gpurun_out/r02c_gpu_tests.log
timeout 420 python bench.py > gpurun_out/r02c_bench_line.json 2> gpurun_out/r02c_bench_line.err; tail -c 300 gpurun_out/r02c_bench_line.err
python - <<'P'
import json
try:
    d = json.load(open('gpurun_out/r02c_bench_line.json'))
    print({k: d[k] for k in ('value', 'ms_per_step', 'kernel_ms')}, 'frac', d['roofline']['frac'], 'e2e ms', d['e2e']['ms_per_step'],
          'verdicts-only ms', d.get('e2e_verdicts_only', {}).get('ms_per_step'), 'tap ms', d.get('transcript_tap', {}).get('ms'), 'cpu', d.get('cpu_baseline', {}).get('value'))
except Exception as ex:
    print('no bench line:', ex)
P
timeout 200 python tools/exp_relay_variants.py 2>&1 | grep -v "templates:" | cut -c1-330 | tee gpurun_out/r02c_variants.log
