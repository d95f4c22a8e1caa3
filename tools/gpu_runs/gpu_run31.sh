#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 3 gpurun_out/$name.log; }
run tests 900 python -m pytest tests -m gpu -q --maxfail=20 --timeout 300
run smoke 300 python __graft_entry__.py --smoke
run bench 900 python bench.py --steps 5 --warmup 3
grep -h '"metric"' gpurun_out/bench.log > gpurun_out/bench_line.json
run bench_ref 600 python bench.py --impl reference --steps 2 --warmup 1
