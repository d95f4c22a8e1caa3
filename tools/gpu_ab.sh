mkdir -p gpurun_out
for wn in 1 0 1; do
  SELFRECON_B200_FUSED_WN=$wn timeout 200 python bench.py --steps 12 --warmup 3 > gpurun_out/bench_wn$wn.log 2>&1
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_wn$wn.log") if l.startswith("{")][-1])
t=d["train"]
print("fused_wn=$wn value", round(d["value"]), "train ms", round(t["ms_per_step"],2), "fwd", round(t["ms_forward_incl_trace"],2), "bwd", round(t["ms_backward"],2), "prop", round(t["ms_propagate"],2), "opt", round(t["ms_optimizer"],2))
PY
done
