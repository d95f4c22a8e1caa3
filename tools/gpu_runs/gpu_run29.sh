#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 3 gpurun_out/$name.log; }
run tests_trace 600 python -m pytest tests -m gpu -q -k "trace or optim or smoke or tc_" --maxfail=10 --timeout 200
run trace_grouped 200 python tools/trace_bench.py
SELFRECON_B200_TC_GROUPED=0 run trace_single 200 python tools/trace_bench.py
