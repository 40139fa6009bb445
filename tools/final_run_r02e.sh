#!/bin/bash
# Last GPU call of the round: full parity suite on the final sources, headline bench line, the step's variants (OpenAI-shaped probe).
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r02d_gpu_tests.log 2>&1; tail -n 2 gpurun_out/r02d_gpu_tests.log
timeout 300 python bench.py > gpurun_out/r02d_bench_line.json 2> gpurun_out/r02d_bench_line.err; tail -c 200 gpurun_out/r02d_bench_line.err
python - <<'P'
import json
try:
    d = json.load(open('gpurun_out/r02d_bench_line.json'))
    print({k: d[k] for k in ('value', 'ms_per_step')}, d['kernel_ms']['step_chained'], 'frac', d['roofline']['frac'], 'e2e ms', d['e2e']['ms_per_step'],
          'vo', d.get('e2e_verdicts_only', {}).get('ms_per_step'), 'tap', d.get('transcript_tap', {}).get('ms'), 'cpu', d.get('cpu_baseline', {}).get('value'))
except Exception as ex:
    print('no bench line:', ex)
P
timeout 120 python tools/exp_relay_variants.py 2>&1 | grep -v "templates:" | cut -c1-330 | tee gpurun_out/r02d_variants.log
