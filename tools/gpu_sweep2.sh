mkdir -p gpurun_out
timeout 200 python -X faulthandler -m pytest tests/test_gpu_round2.py -x -q -s -k "whole_sweep" -o faulthandler_timeout=100 > gpurun_out/sweep_test.log 2>&1; echo "exit $?" >> gpurun_out/sweep_test.log
grep -v "Warn\|WeightNorm" gpurun_out/sweep_test.log | tail -4
timeout 200 python tools/sweep_bench.py 50333 18944 6144 > gpurun_out/sweep_bench.log 2>&1; grep "M=" gpurun_out/sweep_bench.log
