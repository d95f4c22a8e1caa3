#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 2 gpurun_out/$name.log; }
export SELFRECON_B200_GRAPHS=0
run ncu_list 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_d.csv python tools/profile_step.py
run ncu_tc 900 ncu --set full --clock-control none --import-source on -k regex:tc_layer_pair_kernel -s 60 -c 2 -o gpurun_out/prof_tc_d -f python tools/profile_step.py
run ncu_mc 600 ncu --set full --clock-control none --import-source on -k regex:mc_ -c 4 -o gpurun_out/prof_mc_d -f python tools/profile_step.py --mc-only
