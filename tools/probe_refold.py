"""Where does a ray-part call spend its time right after the weights changed (fold / pack / eager trace / shade)?"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def t(fn, n=1):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3 / n, r


def main():
    from selfreconcode_b200 import ops
    dev = torch.device("cuda:0")
    sc = bench.build_scene(dev, 0)
    R = sc["rays"]
    rays, init, bi = R["rays"].to(dev), R["init_pts"].to(dev), R["batch_inds"].to(dev)
    for _ in range(3):
        bench.ray_part(sc, rays, init, bi)
    print("steady ray_part ms: %.2f" % t(lambda: bench.ray_part(sc, rays, init, bi), 3)[0])
    sdf, comp, rn = sc["sdf"], sc["comp"], sc["rn"]
    tr, sk = comp.defs
    for rep in range(2):
        for m in (sdf, tr, rn):
            with torch.no_grad():
                next(iter(m.parameters())).add_(0.0)      # bumps Tensor._version like optimizer.step(): refold in place
        ms_fold, nets = t(lambda: (sdf.fused_sdf_only(), sdf.fused(), tr.fused(bench.RATIO), rn.fused(bench.RATIO)))
        ms_pack, _ = t(lambda: [ops.tc_net(n) for n in nets])
        lbs = sk.lbs_state()
        ms_pose, _ = t(lambda: lbs.set_pose(sc["conds"][1][0], sc["conds"][1][1]))
        nets[0].set_pe_weights([1.0] * 6)
        ms_trace, out = t(lambda: ops.trace_surface_points(nets[0], nets[2], lbs, sc["cam"]["cam_pos"], rays, init, bi,
                                                           sc["conds"][0], 5e-5, sc["ang"], 3.05, 1.0, 10,
                                                           return_counters=True))
        ms_trace2, _ = t(lambda: ops.trace_surface_points(nets[0], nets[2], lbs, sc["cam"]["cam_pos"], rays, init, bi,
                                                          sc["conds"][0], 5e-5, sc["ang"], 3.05, 1.0, 10,
                                                          return_counters=True))
        ms_shade, _ = t(lambda: ops.shade_and_render_tc(nets[1], nets[2], lbs, nets[3], out[0], rays, bi, sc["conds"][0]))
        print("rep %d: fold %.2f  pack %.2f  pose %.2f  trace(eager) %.2f  trace(capture) %.2f  shade %.2f ms"
              % (rep, ms_fold, ms_pack, ms_pose, ms_trace, ms_trace2, ms_shade))
    os.environ["X"] = "1"
    ops.GRAPHS_ENABLED = False
    for m in (sdf, tr, rn):
        with torch.no_grad():
            next(iter(m.parameters())).add_(0.0)
    nets = (sdf.fused_sdf_only(), sdf.fused(), tr.fused(bench.RATIO), rn.fused(bench.RATIO))
    nets[0].set_pe_weights([1.0] * 6)
    lbs = sk.lbs_state()
    f = lambda: ops.trace_surface_points(nets[0], nets[2], lbs, sc["cam"]["cam_pos"], rays, init, bi, sc["conds"][0], 5e-5,
                                         sc["ang"], 3.05, 1.0, 10)
    print("eager trace (no graphs): first %.2f, then %.2f ms" % (t(f)[0], t(f, 3)[0]))


if __name__ == "__main__":
    main()
