#!/bin/bash
# full GPU suite + the new bench configurations
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gpu_tests.log
tail -15 gpurun_out/r02_gpu_tests.log
timeout 600 python bench.py --config c4 --steps 5 --warmup 3 > gpurun_out/r02_bench_c4_n1.json 2> gpurun_out/r02_bench_c4_n1.err; echo "c4 rc=$?"; tail -c 1500 gpurun_out/r02_bench_c4_n1.json; tail -5 gpurun_out/r02_bench_c4_n1.err
timeout 600 python bench.py --config c5 --steps 5 --warmup 3 > gpurun_out/r02_bench_c5_n1.json 2> gpurun_out/r02_bench_c5_n1.err; echo "c5 rc=$?"; tail -c 2500 gpurun_out/r02_bench_c5_n1.json; tail -5 gpurun_out/r02_bench_c5_n1.err
