#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 6 gpurun_out/$name.log; }
run optim_tests 600 python -m pytest tests/test_optim_propagate.py tests/test_optim_step_gpu.py -m gpu -q --maxfail=20
for v in base noepi nomma noload noloadw split nomma_noepi noload_noepi; do
  if [ $v = base ]; then unset SELFRECON_B200_LIB; else export SELFRECON_B200_LIB=$PWD/selfreconcode_b200/lib/variants/libselfrecon_b200_$v.so; fi
  run diag_$v 120 python tools/tc_bench.py
done
