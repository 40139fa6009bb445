#!/bin/bash
# Evidence for the final sources (2-deep ring): ncu launch list of the bench command, one full capture of k_relay2 (DRAM traffic),
# and the C4 / C5 lines.
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02c_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_ncu_l.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_relay2 -s 3 -c 1 -f -o gpurun_out/r02c_relay2 python tools/exp_one.py > /dev/null 2>&1
timeout 300 python bench.py --config c4 --steps 5 --warmup 3 > gpurun_out/r02c_bench_c4_n1.json 2> gpurun_out/r02c_bench_c4_n1.err; cut -c1-260 gpurun_out/r02c_bench_c4_n1.json
timeout 300 python bench.py --config c5 --steps 5 --warmup 3 > gpurun_out/r02c_bench_c5_n1.json 2> gpurun_out/r02c_bench_c5_n1.err; cut -c1-260 gpurun_out/r02c_bench_c5_n1.json
ls -la gpurun_out/r02c_*
