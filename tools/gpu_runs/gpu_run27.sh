#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 3 gpurun_out/$name.log; }
run tests_interp 600 python -m pytest tests -m gpu -q -k "interp or seg3d or smoke" --maxfail=10 --timeout 300
run ref_ab 300 python tools/ref_ab_bench.py
