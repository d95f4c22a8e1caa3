mkdir -p gpurun_out
timeout 300 python -X faulthandler -m pytest tests/test_gpu_round2.py -x -q -s -k "whole_sweep" -o faulthandler_timeout=100 > gpurun_out/sweep_test.log 2>&1; echo "exit $?" >> gpurun_out/sweep_test.log
grep -v "Warn\|WeightNorm" gpurun_out/sweep_test.log | tail -25
if grep -q "1 passed" gpurun_out/sweep_test.log; then
  timeout 300 python bench.py --steps 10 --warmup 3 --no-train > gpurun_out/bench_sweep.log 2>&1
  SELFRECON_B200_TC_SWEEP=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-train > gpurun_out/bench_nosweep.log 2>&1
  python - <<'PY'
import json
for f in ("gpurun_out/bench_sweep.log","gpurun_out/bench_nosweep.log"):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, d["value"], d["ms_ray_part"], d["ms_mc_part"], d["roofline"]["trace"]["ms"], d["e2e"]["value"], d["gpu_launches"])
    except Exception as e: print(f, "fail", e)
PY
fi
