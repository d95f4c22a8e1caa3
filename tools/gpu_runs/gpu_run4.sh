#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 6 gpurun_out/$name.log; }
run tests 600 python -m pytest tests -m gpu -q --maxfail=20
for u in 2 4 8; do
  SELFRECON_B200_LIB=$PWD/selfreconcode_b200/lib/variants/libselfrecon_b200_u$u.so run mb_u$u 300 python tools/microbench.py
done
grep -h '"lib"' gpurun_out/mb_u*.log > gpurun_out/microbench.jsonl
