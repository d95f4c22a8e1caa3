#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 8 gpurun_out/$name.log; }
run tc_test 400 python -m pytest tests -m gpu -q --maxfail=5 -k "tc_ or trace"
run tc_bench 180 python tools/tc_bench.py
run bench 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline
grep -h '"metric"' gpurun_out/bench.log > gpurun_out/bench_line.json
