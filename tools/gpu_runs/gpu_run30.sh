#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/tc_sweep.py 2>/dev/null | tail -1
