#!/bin/bash
# Second GPU pass: full gpu test suite, bench, ncu launch list + full captures of the top kernels.
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 12 gpurun_out/$name.log; }
run tests 600 python -m pytest tests -m gpu -q --maxfail=20
run bench 900 python bench.py --steps 3 --warmup 3
grep -h '"metric"' gpurun_out/bench.log > gpurun_out/bench_line.json
run bench_ref 300 python bench.py --impl reference --steps 1 --warmup 1
# every launch with its device time (cold, serialised: compare shares)
run ncu_list 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py
# full capture of the dominant kernels
run ncu_trace 900 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 2 -c 2 -o gpurun_out/prof_trace -f python tools/profile_step.py
run ncu_mc 600 ncu --set full --clock-control none --import-source on -k regex:mc_ -c 6 -o gpurun_out/prof_mc -f python tools/profile_step.py --mc-only
run ncu_sdf 600 ncu --set full --clock-control none --import-source on -k regex:sdf_kernel -s 3 -c 1 -o gpurun_out/prof_sdf -f python tools/profile_step.py --mc-only
