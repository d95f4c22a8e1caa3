"""Whole-sweep kernel against one launch per layer on the full-size SDF (forward sweep + reverse sweep, 1 row per
point) at the bench's ray count and at smaller batches; CUDA events over back-to-back repetitions, cold operands
between variants.  `python tools/sweep_bench.py [M ...]`"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from selfreconcode_b200 import _lib, ops, synth
    lib = _lib.load()
    dev = torch.device("cuda:0")
    sdf = synth.make_sdf().to(dev)
    net = sdf.fused_sdf_only()
    net.set_pe_weights([1.0] * 6)
    tcn = ops.tc_net(net)
    Ms = [int(a) for a in sys.argv[1:]] or [50333, 18944, 6144]
    for P in Ms:
        B = ops._TcTraceBuffers(dev, P, net, None)
        pts = (torch.rand(P, 3, device=dev) - 0.5)
        pw = (C.c_float * 16)(*[net.desc.pe_w[i] for i in range(16)])
        ops.check(lib.sr_tc_embed(ops._p(pts), P, net.desc.multires, pw, 1, None, None, 0, 0, ops._p(B.emb_s), B.ld_s,
                                  None, None, ops._stream()), "embed")
        B.cot_s.normal_()
        flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)

        def fwd():
            ops._tc_forward_sweep(lib, net, tcn, B.emb_s, B.ld_s, B.A_in, B.acts_s, P, None, B.f)

        def bwd():
            ops._tc_backward_sweep(lib, net, tcn, B.cot_s, B.A, B.acts_s, P, None, B.gs, B.gskip, net.desc.d_in)

        res = {}
        for name, sweep, dbg in (("per-layer", False, 0), ("sweep", True, 0), ("sweep, no epilogue fence", True, 1)):
            ops.TC_SWEEP = sweep
            lib.sr_tc_debug_sweep_flags(dbg)
            for fn, tag in ((fwd, "fwd"), (bwd, "bwd")):
                for _ in range(3):
                    fn()
                flush.zero_()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                res[(name, tag)] = e0.elapsed_time(e1) / 10
            lib.sr_tc_debug_sweep_flags(0)
        ops.TC_SWEEP = True
        for k, v in res.items():
            print("M=%6d %-26s %s sweep of 9 layers: %.1f us (%.1f us per layer incl. pack)" % (P, k[0], k[1], v * 1e3, v * 1e3 / 9))


if __name__ == "__main__":
    main()
