#!/bin/bash
# one full ncu capture each of k_relay2 and k_commit2 on the C3 step (source-level: built with -lineinfo)
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_relay2 -s 3 -c 1 -f -o gpurun_out/r02b_relay2 python tools/exp_one.py > gpurun_out/r02b_ncu_relay2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_commit2 -s 3 -c 1 -f -o gpurun_out/r02b_commit2 python tools/exp_one.py > gpurun_out/r02b_ncu_commit2.log 2>&1
ls -la gpurun_out/*.ncu-rep
