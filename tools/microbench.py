"""Kernel-level timings on the bench scene (CUDA events, L2 flushed, median of N).  Tuning aid."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from selfreconcode_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
sc = bench.build_scene(dev, 0)
R = sc["rays"]
rays, init, bi = R["rays"].to(dev), R["init_pts"].to(dev), R["batch_inds"].to(dev)
P = rays.shape[0]
sdf, comp, rn = sc["sdf"], sc["comp"], sc["rn"]
tr, sk = comp.defs
RATIO = bench.RATIO
full = sdf.fused()
sdf_only = sdf.fused_sdf_only()
dnet = tr.fused(RATIO)
lbs = sk.lbs_state()
lbs.set_pose(sc["conds"][1][0], sc["conds"][1][1])
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
big = torch.rand(400000, 3, device=dev) * 2 - 1


def timeit(fn, n=5):
    ts = []
    for _ in range(n + 1):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts[1:]))


F_S, F_D, F_R = bench.F_S, bench.F_D, bench.F_R
out = {"lib": os.environ.get("SELFRECON_B200_LIB", "default"), "rays": P}
t = timeit(lambda: ops.sdf_forward(sdf_only, big, False, 0))
out["sdf_T0_400k_ms"] = t
out["sdf_T0_TFLOPs"] = 400000 * F_S / t / 1e9
t = timeit(lambda: ops.sdf_forward(full, init, True, 256))
out["sdf_T3_feat_ms"] = t
out["sdf_T3_exec_TFLOPs"] = P * 4 * F_S / t / 1e9
t = timeit(lambda: ops.deform_forward(dnet, lbs, init, bi, sc["conds"][0], True))
out["deform_T3_ms"] = t
out["deform_T3_exec_TFLOPs"] = P * 4 * F_D / t / 1e9
for mode in ("forward", "reverse"):
    t = timeit(lambda: ops.trace_surface_points(sdf_only, dnet, lbs, sc["cam"]["cam_pos"], rays, init, bi,
                                                sc["conds"][0], 5e-5, sc["ang"], 3.05, 1.0, 10, mode=mode), 3)
    out["trace_%s_ms" % mode] = t
_, _, cnt = ops.trace_surface_points(sdf_only, dnet, lbs, sc["cam"]["cam_pos"], rays, init, bi, sc["conds"][0],
                                     5e-5, sc["ang"], 3.05, 1.0, 10, return_counters=True)
it = sum(cnt.cpu().tolist()[1:11])
out["trace_alg_TFLOPs_reverse"] = (P + 3 * it) * (F_S + F_D) / out["trace_reverse_ms"] / 1e9
out["trace_alg_TFLOPs_forward"] = (P + 3 * it) * (F_S + F_D) / out["trace_forward_ms"] / 1e9
grid = torch.rand(257, 257, 257, device=dev) - 0.5
xx = torch.linspace(-1, 1, 257, device=dev)
g3 = (xx.view(-1, 1, 1) ** 2 + xx.view(1, -1, 1) ** 2 + xx.view(1, 1, -1) ** 2).sqrt() - 0.6
t = timeit(lambda: ops.marching_cubes(g3.contiguous(), 1, 1, 1, 0, 0, 0, 0.0))
out["mc_257_ms"] = t
print(json.dumps(out))
