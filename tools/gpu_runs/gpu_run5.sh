#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 6 gpurun_out/$name.log; }
run tests 600 python -m pytest tests -m gpu -q --maxfail=20
SELFRECON_B200_LIB=$PWD/selfreconcode_b200/lib/variants/libselfrecon_b200_w16.so run tests_w16 600 python -m pytest tests -m gpu -q --maxfail=20 -k "sdf or deform or render or trace or cardinal or seg3d"
for v in w8 w16 w8slow; do
  SELFRECON_B200_LIB=$PWD/selfreconcode_b200/lib/variants/libselfrecon_b200_$v.so run mb_$v 300 python tools/microbench.py
done
grep -h '"lib"' gpurun_out/mb_w*.log > gpurun_out/microbench.jsonl
