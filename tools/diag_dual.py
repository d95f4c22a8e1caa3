"""Diagnostic: tensor-core tracer with / without the second stream, eager and graph-replayed, against the fp32 engine.
    python tools/diag_dual.py <dual 0|1> <calls>
Prints position / mask statistics per call; dumps the Python stack if a call blocks for 60 s."""
import faulthandler
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    dual, calls = int(sys.argv[1]), int(sys.argv[2])
    faulthandler.dump_traceback_later(60, exit=True)
    from selfreconcode_b200 import ops, synth
    ops.TC_DUAL_STREAM = bool(dual)
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    sdf = synth.make_sdf().to(dev)
    tr = synth.make_translator().to(dev)
    sk = synth.make_skinner(resolution=(33, 57, 17)).to(dev)
    comp = synth.CompositeDeformer([tr, sk]).to(dev)
    poses, trans, dcond = [t.to(dev) for t in synth.make_frame_params(5, 2)]
    ratio = {"sdfRatio": 1.0, "deformerRatio": 1.0, "renderRatio": 1.0}
    cam = synth.camera(128, 128)
    Pt = 16384
    dirs = torch.nn.functional.normalize(torch.randn(Pt, 3, generator=g), dim=1)
    dirs[:, 2] = -dirs[:, 2].abs()
    start = (0.6 * dirs + 2e-3 * torch.randn(Pt, 3, generator=g)).to(dev)
    bt = torch.randint(0, 2, (Pt,), generator=g).to(dev)
    sdf_only = sdf.fused_sdf_only()
    sdf_only.set_pe_weights([1.0] * 6)
    dnet = tr.fused(ratio)
    lbs = sk.lbs_state()
    lbs.set_pose(poses, trans)
    with torch.no_grad():
        dst = comp.forward_fused(start, [dcond, [poses, trans]], bt, ratio)[0]
    cpos = cam["cam_pos"].to(dev)
    rays = torch.nn.functional.normalize(dst - cpos.view(1, 3), dim=1)
    ref = ops.trace_surface_points(sdf_only, dnet, lbs, cpos, rays, start, bt, dcond, 5e-5, 0.05, 3.05, 1.0, 5,
                                   mode="reverse")
    torch.cuda.synchronize()
    print("fp32 engine: %d converged" % int(ref[1].sum()), flush=True)
    for c in range(calls):
        out = ops.trace_surface_points(sdf_only, dnet, lbs, cpos, rays, start, bt, dcond, 5e-5, 0.05, 3.05, 1.0, 5,
                                       mode="tc")
        torch.cuda.synchronize()
        dp = (out[0] - ref[0]).abs().max(dim=1).values
        both = out[1] & ref[1]
        print("dual %d call %d: max|dp| all %.3e, on rays converged in both %.3e (%d rays), >7e-5: %d, mask diffs %d"
              % (dual, c, dp.max().item(), dp[both].max().item(), int(both.sum()), int((dp > 7e-5).sum()),
                 int((out[1] != ref[1]).sum())), flush=True)
    faulthandler.cancel_dump_traceback_later()
    print("done", flush=True)


if __name__ == "__main__":
    main()
