#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 5 gpurun_out/$name.log; }
run pair_test 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tc_" --timeout 60
run pair_bench 120 python tools/tc_bench.py
SELFRECON_B200_TC_PAIR=0 run single_bench 120 python tools/tc_bench.py
