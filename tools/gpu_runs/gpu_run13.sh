#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 2 gpurun_out/$name.log; }
run mc_breakdown 300 python tools/mc_breakdown.py
