#!/bin/bash
# Round-2 end measurements on one B200 (run through gpurun): smoke, GPU tests, bench lines (C3 headline + reference arm, C4, C5),
# ncu launch list of the bench command, one full capture of k_relay2 and k_commit2.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_final_gpu_tests.log 2>&1; tail -n 3 gpurun_out/r02_final_gpu_tests.log
timeout 900 python bench.py > gpurun_out/r02_bench_line.json 2> gpurun_out/r02_bench_line.err; tail -c 400 gpurun_out/r02_bench_line.err; cut -c1-600 gpurun_out/r02_bench_line.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; cut -c1-400 gpurun_out/r02_bench_reference_arm.json
timeout 600 python bench.py --config c4 --steps 5 --warmup 3 > gpurun_out/r02_bench_c4_n1.json 2> gpurun_out/r02_bench_c4_n1.err; cut -c1-200 gpurun_out/r02_bench_c4_n1.json
timeout 600 python bench.py --config c5 --steps 5 --warmup 3 > gpurun_out/r02_bench_c5_n1.json 2> gpurun_out/r02_bench_c5_n1.err; cut -c1-200 gpurun_out/r02_bench_c5_n1.json
timeout 600 python bench.py --impl reference --config c4 --steps 1 --warmup 0 > gpurun_out/r02_bench_reference_c4.json 2> gpurun_out/r02_bench_reference_c4.err; cut -c1-300 gpurun_out/r02_bench_reference_c4.json
timeout 600 python bench.py --impl reference --config c5 > gpurun_out/r02_bench_reference_c5.json 2> gpurun_out/r02_bench_reference_c5.err; cut -c1-300 gpurun_out/r02_bench_reference_c5.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_l.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_relay2 -s 3 -c 1 -f -o gpurun_out/r02_relay2_final python tools/exp_one.py > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_commit2 -s 3 -c 1 -f -o gpurun_out/r02_commit2_final python tools/exp_one.py > /dev/null 2>&1
timeout 300 python tools/exp_relay_variants.py > gpurun_out/r02_variants.log 2>&1; tail -n 8 gpurun_out/r02_variants.log | cut -c1-400
ls -la gpurun_out/r02_*final* gpurun_out/r02_launches.csv
