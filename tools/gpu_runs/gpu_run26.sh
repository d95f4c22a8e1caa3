#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 3 gpurun_out/$name.log; }
run tests 900 python -m pytest tests -m gpu -q --maxfail=20 --timeout 300
run ref_ab 300 python tools/ref_ab_bench.py
