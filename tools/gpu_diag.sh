mkdir -p gpurun_out
for cfg in "0 1" "0 4" "1 1" "1 4"; do
  set -- $cfg
  echo "=== dual $1 calls $2" >> gpurun_out/diag.log
  timeout 150 python tools/diag_dual.py $1 $2 >> gpurun_out/diag.log 2>&1; echo "exit $?" >> gpurun_out/diag.log
done
echo "=== golden step dual 0" >> gpurun_out/diag.log
SELFRECON_B200_TC_DUAL_STREAM=0 timeout 240 python -X faulthandler -m pytest tests/test_gpu_forward_golden.py -x -q -o faulthandler_timeout=120 >> gpurun_out/diag.log 2>&1; echo "exit $?" >> gpurun_out/diag.log
grep -v "Warn\|WeightNorm" gpurun_out/diag.log | tail -60
