#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 3 gpurun_out/$name.log; }
run ncu_list 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c.csv python tools/profile_step.py
run ncu_tc 900 ncu --set full --clock-control none --import-source on -k regex:tc_layer_kernel -s 60 -c 2 -o gpurun_out/prof_tc_c -f python tools/profile_step.py
run ncu_mc 600 ncu --set full --clock-control none --import-source on -k regex:mc_ -c 3 -o gpurun_out/prof_mc_c -f python tools/profile_step.py --mc-only
run bench_ref 600 python bench.py --impl reference --steps 2 --warmup 1
