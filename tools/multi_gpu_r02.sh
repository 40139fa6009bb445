#!/bin/bash
# bench lines at N GPUs (run through gpurun --gpus N): C3 (weak), C4 (strong), C5 (strong)
N=$1; shift
mkdir -p gpurun_out
for c in "$@"; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 --config $c > gpurun_out/r02_bench_${c}_n$N.json 2> gpurun_out/r02_bench_${c}_n$N.err
  echo "$c N=$N rc=$?"; cut -c1-260 gpurun_out/r02_bench_${c}_n$N.json; grep -i "error\|assert" gpurun_out/r02_bench_${c}_n$N.err | head -3
done
