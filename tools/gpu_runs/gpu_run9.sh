#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 6 gpurun_out/$name.log; }
run tests 900 python -m pytest tests -m gpu -q --maxfail=20
for v in base noepi nomma noload; do
  if [ $v = base ]; then unset SELFRECON_B200_LIB; else export SELFRECON_B200_LIB=$PWD/selfreconcode_b200/lib/variants/libselfrecon_b200_$v.so; fi
  run diag_$v 120 python tools/tc_bench.py
done
unset SELFRECON_B200_LIB
run bench 900 python bench.py --steps 5 --warmup 3
grep -h '"metric"' gpurun_out/bench.log > gpurun_out/bench_line.json
