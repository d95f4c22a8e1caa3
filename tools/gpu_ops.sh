mkdir -p gpurun_out
timeout 300 python tools/step_ops.py > gpurun_out/step_ops.log 2>&1
tail -5 gpurun_out/step_ops.log
