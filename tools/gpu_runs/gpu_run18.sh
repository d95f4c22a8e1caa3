#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 1 gpurun_out/$name.log; }
SELFRECON_B200_TC_TAILSPLIT=1 run split1 120 python tools/tc_bench.py
SELFRECON_B200_TC_TAILSPLIT=0 run split0 120 python tools/tc_bench.py
SELFRECON_B200_TC_TAILSPLIT=1 run split1b 120 python tools/tc_bench.py
SELFRECON_B200_TC_TAILSPLIT=0 run split0b 120 python tools/tc_bench.py
