mkdir -p gpurun_out
echo "== 8 epilogue warps" > gpurun_out/sweep_bench2.log
timeout 200 python tools/sweep_bench.py 50333 262144 >> gpurun_out/sweep_bench2.log 2>&1
echo "== 16 epilogue warps" >> gpurun_out/sweep_bench2.log
SELFRECON_B200_LIB=$PWD/selfreconcode_b200/lib/variants/libselfrecon_b200_e16.so timeout 200 python tools/sweep_bench.py 50333 262144 >> gpurun_out/sweep_bench2.log 2>&1
SELFRECON_B200_LIB=$PWD/selfreconcode_b200/lib/variants/libselfrecon_b200_e16.so timeout 200 python -X faulthandler -m pytest tests/test_gpu_round2.py -x -q -s -k "whole_sweep" -o faulthandler_timeout=100 > gpurun_out/sweep_test16.log 2>&1
grep "M=\|==" gpurun_out/sweep_bench2.log | grep -v "no epilogue"; tail -2 gpurun_out/sweep_test16.log
