#!/bin/bash
# Round-end measurements on one B200 (run through gpurun): smoke, GPU tests, bench line, ncu launch list, one full capture.
V=${1:-v17}
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 4
timeout 600 python bench.py > gpurun_out/bench_$V.json 2> gpurun_out/bench_$V.err; tail -c 600 gpurun_out/bench_$V.err; cut -c1-900 gpurun_out/bench_$V.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$V.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l_$V.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_relay$ --launch-skip 6 --launch-count 1 -f -o gpurun_out/prof_relay_$V python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f_$V.log 2>&1
tail -n 2 gpurun_out/ncu_f_$V.log | cut -c1-300
