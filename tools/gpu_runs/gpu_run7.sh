#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 8 gpurun_out/$name.log; }
run tests 600 python -m pytest tests -m gpu -q --maxfail=20
run smoke 240 python __graft_entry__.py --smoke
run bench 900 python bench.py --steps 5 --warmup 3
grep -h '"metric"' gpurun_out/bench.log > gpurun_out/bench_line.json
run ncu_list 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_b.csv python tools/profile_step.py
run ncu_tc 900 ncu --set full --clock-control none --import-source on -k regex:tc_layer_kernel -s 40 -c 3 -o gpurun_out/prof_tc -f python tools/profile_step.py
run ncu_mc 600 ncu --set full --clock-control none --import-source on -k regex:mc_ -c 3 -o gpurun_out/prof_mc_b -f python tools/profile_step.py --mc-only
