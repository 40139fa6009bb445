#!/bin/bash
# usage: tools/run_variants.sh name1 name2 ...   (tools/lib_<name>.so; "main" = the built library)
# Runs tools/exp_relay_variants.py once per library variant (perf probes only; the shipped library is restored).
set -u
N=llmapigateway_b200/_native/libllmgw_b200.so
cp $N /tmp/lib_main.so
for v in "$@"; do
  if [ "$v" = main ]; then cp /tmp/lib_main.so $N; else cp tools/lib_$v.so $N; fi
  echo "VARIANT $v"
  timeout 300 python tools/exp_relay_variants.py 2>&1 | grep -v "template cache" | cut -c1-200
done
cp /tmp/lib_main.so $N
