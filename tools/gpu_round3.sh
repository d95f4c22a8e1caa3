mkdir -p gpurun_out
timeout 600 python -X faulthandler -m pytest tests -m gpu -x -q -o faulthandler_timeout=150 > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
timeout 240 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/smoke.log 2>&1
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv python tools/train_step_once.py > gpurun_out/train_once.log 2>&1
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-train > gpurun_out/bench_ncu.log 2>&1
grep -v "Warn\|WeightNorm" gpurun_out/smoke.log | tail -2; tail -3 gpurun_out/pytest.log
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench.log") if l.startswith("{")][-1])
print("value", d["value"], "ray", d["ms_ray_part"], "mc", d["ms_mc_part"], "trace", d["roofline"]["trace"]["ms"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], d["roofline"]["ms_per_launch"])
t=d.get("train") or {}
print("train", t.get("ms_per_step"), t.get("value"), t.get("ms_forward_incl_trace"), t.get("ms_backward"), t.get("ms_propagate"), t.get("gpu_launches_own_kernels_per_step"))
PY
wc -l gpurun_out/launches.csv gpurun_out/train_launches.csv
