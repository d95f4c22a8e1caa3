mkdir -p gpurun_out
V=$PWD/selfreconcode_b200/lib/variants/libselfrecon_b200_r16.so
echo "== main" > gpurun_out/r16.log
timeout 100 python tools/sweep_bench.py 50333 2>&1 | grep "per-layer" >> gpurun_out/r16.log
timeout 100 python tools/prof_train_kernels.py 2>&1 | tail -2 >> gpurun_out/r16.log
echo "== reverse epilogues on 16 warps" >> gpurun_out/r16.log
SELFRECON_B200_LIB=$V timeout 100 python tools/sweep_bench.py 50333 2>&1 | grep "per-layer" >> gpurun_out/r16.log
SELFRECON_B200_LIB=$V timeout 100 python tools/prof_train_kernels.py 2>&1 | tail -2 >> gpurun_out/r16.log
SELFRECON_B200_LIB=$V timeout 300 python -m pytest tests/test_gpu_train.py tests/test_gpu_round2.py -x -q > gpurun_out/r16_test.log 2>&1
cat gpurun_out/r16.log; tail -2 gpurun_out/r16_test.log
